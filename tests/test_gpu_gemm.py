"""The tcgen05/TMA 3xTF32 GEMM against float64 matmul and against the FFMA GEMM, all four operand-major
combinations, overlapping-row (frame) operands, ragged M / K, split-K."""
import numpy as np
import pytest
import torch

from oracle import st_oracle as O

pytestmark = pytest.mark.gpu


def _engine():
    from signaltrain_b200.engine import Engine, Geometry
    return Engine(Geometry(1, 4, 4), "cuda:0")


def _split(x):
    """tf32 pair: hi keeps the top 19 bits (10-bit mantissa), lo = tf32(x - hi).  hi + lo == x to 2^-22."""
    xi = x.view(np.uint32)
    hi = ((xi + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    r = (x - hi).astype(np.float32)
    lo = ((r.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, lo


def _operand(rng, rows, cols, ld):
    """Row-major (rows, cols) view with leading dim ld (ld < cols => overlapping rows) over one flat buffer."""
    n = (rows - 1) * ld + cols
    flat = (rng.standard_normal(n) * rng.uniform(0.1, 2.0)).astype(np.float32)
    hi, lo = _split(flat)
    view = np.lib.stride_tricks.as_strided((hi.astype(np.float64) + lo), shape=(rows, cols), strides=(ld * 8, 8))
    pad = np.zeros(1024, np.float32)          # TMA boxes may start inside and run past the last row: keep memory mapped
    return torch.from_numpy(np.concatenate([hi, pad])).cuda(), torch.from_numpy(np.concatenate([lo, pad])).cuda(), view


CASES = [
    # (a_mn, b_mn, M, N, K, a_ld, b_ld, splits)     ld=None -> dense
    (0, 0, 300, 1056, 1024, 384, None, 1),      # analysis: frames (overlapping rows, hop 384) x wcat^T
    (0, 0, 5400, 1056, 1024, 384, None, 1),     # full size, persistent over 258 tiles
    (0, 1, 2200, 1024, 1056, None, None, 1),    # synthesis: ri x sfold
    (1, 1, 1056, 1024, 333, None, 384, 1),      # weight gradient, ragged K, B rows overlap
    (1, 1, 1056, 1024, 2200, None, 384, 4),     # split-K
    (1, 0, 256, 512, 64, None, None, 1),
]


@pytest.mark.parametrize("a_mn,b_mn,M,N,K,a_ld,b_ld,splits", CASES)
def test_tcgen05_gemm(a_mn, b_mn, M, N, K, a_ld, b_ld, splits):
    eng = _engine()
    rng = np.random.RandomState(M + N + K)
    ar, ac = (K, M) if a_mn else (M, K)
    br, bc = (K, N) if b_mn else (N, K)
    a_ld = a_ld or ac
    b_ld = b_ld or bc
    ah, al, A = _operand(rng, ar, ac, a_ld)
    bh, bl, Bm = _operand(rng, br, bc, b_ld)
    A2 = A.T if a_mn else A            # (M, K)
    B2 = Bm if b_mn else Bm.T          # (K, N)
    ref = A2 @ B2
    out = eng.debug_gemm(1, a_mn, b_mn, ah, al, a_ld, bh, bl, b_ld, M, N, K, splits).sum(0).cpu().numpy()
    scale = np.abs(ref).max()
    err = np.abs(out - ref).max() / scale
    simt = eng.debug_gemm(0, a_mn, b_mn, ah, al, a_ld, bh, bl, b_ld, M, N, K, splits).sum(0).cpu().numpy()
    err_simt = np.abs(simt - ref).max() / scale
    bad = np.argwhere(~(np.abs(out - ref) < 1e-4 * scale))
    print(f"\nGEMM a_mn={a_mn} b_mn={b_mn} M={M} N={N} K={K}: tc err {err:.3e}  simt err {err_simt:.3e}  bad {len(bad)}/{out.size}"
          f" first bad {bad[:4].tolist()} rows-with-bad {np.unique(bad[:, 0])[:12].tolist()} cols-with-bad {np.unique(bad[:, 1])[:12].tolist()}")
    assert np.isfinite(out).all()
    assert err_simt < 2e-6
    assert err < 3e-5, err              # TMEM accumulation rounds toward zero at every MMA step
    # promoted accumulation (fresh TMEM accumulator per k-block, partials summed in fp32 registers): fp32-level
    acc = eng.debug_gemm(2, a_mn, b_mn, ah, al, a_ld, bh, bl, b_ld, M, N, K, splits).sum(0).cpu().numpy()
    err_acc = np.abs(acc - ref).max() / scale
    print(f"     promoted: err {err_acc:.3e}")
    assert err_acc < 2e-6, err_acc
