"""GPU: the reduced-precision mode (st_set_precision(ST_PRECISION_TF32); BASELINE configs 3 and 5 are its class: mixed
precision with fp32 master weights, optimiser state, loss and atan2) against the float64 oracle and against the oracle's
emulation of it (oracle.operand_rounding: every contraction operand rounded to TF32).

The reference reaches reduced precision only through apex (train.py:133-136,169,184), absent here, so there is no golden
for it.  Stated bounds (inputs O(1), goldens of SURVEY section 8d): output waveforms within 2e-3 max-abs of the exact
answer and within 4x (+1e-4) of the error the emulation itself makes; gradients within 3e-2 of each tensor's max-abs (or
twice the emulation's own error where that is larger) and within 6x (+2e-3) of the emulation's error.
The default mode must be untouched by a round trip through the reduced one (bit-identical outputs)."""
import numpy as np
import pytest
import torch

from oracle import st_oracle as O
from tests.conftest import GOLDEN_CASES
from tests.helpers import initial_params, load_case
from tests.test_gpu_parity import _dev_params, _engine, _t

pytestmark = pytest.mark.gpu

WAVE_TOL_TF32 = 2e-3
GRAD_RTOL_TF32 = 3e-2


def _gemm_case(eng, use_tc, a_mn, b_mn, M, N, K, splits, seed):
    rng = np.random.RandomState(seed)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bm = rng.standard_normal((N, K)).astype(np.float32)
    Ah, Bh = O.round_tf32(A), O.round_tf32(Bm)
    a = _t(Ah.T if a_mn else Ah)
    b = _t(Bh.T if b_mn else Bh)
    za, zb = torch.zeros_like(a), torch.zeros_like(b)
    C = eng.debug_gemm(use_tc, a_mn, b_mn, a, za, a.shape[1], b, zb, b.shape[1], M, N, K, splits)
    return C.sum(0).cpu().numpy(), Ah.astype(np.float64) @ Bh.astype(np.float64).T


@pytest.mark.parametrize("a_mn,b_mn,splits", [(0, 0, 1), (0, 1, 1), (1, 1, 4), (1, 0, 1)])
def test_single_pass_gemm_is_exact_on_tf32_operands(a_mn, b_mn, splits):
    """use_tc=3: hi planes only, one kind::tf32 UMMA per k-step.  On operands that ARE tf32 numbers the products are
    exact, so only fp32 accumulation (TMEM, round toward zero: ~7e-6 relative over K=1024) separates it from float64."""
    d = O.model_dims(1, 4, 4)
    eng = _engine(d)
    M, N, K = 640, 512, 1024
    got, ref = _gemm_case(eng, 3, a_mn, b_mn, M, N, K, splits, seed=a_mn * 2 + b_mn)
    scale = np.abs(ref).max()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 3e-5 * scale, np.abs(got - ref).max() / scale      # same bar as tests/test_gpu_gemm.py
    # and the lo planes are really ignored: garbage there must not change the result
    rng = np.random.RandomState(9)
    A = O.round_tf32(rng.standard_normal((M, K)).astype(np.float32))
    Bm = O.round_tf32(rng.standard_normal((N, K)).astype(np.float32))
    a, b = _t(A.T if a_mn else A), _t(Bm.T if b_mn else Bm)
    junk_a, junk_b = torch.full_like(a, 1e6), torch.full_like(b, -1e6)
    C1 = eng.debug_gemm(3, a_mn, b_mn, a, junk_a, a.shape[1], b, junk_b, b.shape[1], M, N, K, splits).sum(0)
    C0 = eng.debug_gemm(3, a_mn, b_mn, a, torch.zeros_like(a), a.shape[1], b, torch.zeros_like(b), b.shape[1], M, N, K, splits).sum(0)
    assert torch.equal(C0, C1)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_tf32_forward_and_gradients_within_stated_bounds(case):
    g, d = load_case(case)
    P = initial_params(g, d)
    x, y, knobs = g["step0/x"], g["step0/y"].astype(np.float32), g["step0/knobs"]
    sbf = O.scale_by_freq(d.F)
    ref_loss, ref_grads, fw = O.loss_and_grads(d, P, x, y, knobs, sbf)
    with O.operand_rounding("tf32"):
        emu_loss, emu_grads, emu = O.loss_and_grads(d, P, x, y, knobs, sbf)
    emu_err = np.abs(emu["y_hat"] - fw["y_hat"]).max()

    eng = _engine(d)
    params = _dev_params(P, d)
    y32, _, mh32, _ = eng.forward(_t(x), _t(knobs), params)                   # default mode first
    eng.set_precision("tf32")
    assert eng.precision == "tf32"
    y_hat, mag, mag_hat, _ = eng.forward(_t(x), _t(knobs), params)
    err = np.abs(y_hat.cpu().numpy() - fw["y_hat"]).max()
    assert err < WAVE_TOL_TF32, err
    assert err < 4 * emu_err + 1e-4, (err, emu_err)
    assert err > 1e-6, "the reduced mode produced fp32-exact output: the mode switch did nothing"
    np.testing.assert_allclose(mag.cpu().numpy(), fw["mag"], atol=2e-2, rtol=2e-3)
    np.testing.assert_allclose(mag_hat.cpu().numpy(), fw["mag_hat"], atol=max(1e-2, 6 * np.abs(emu["mag_hat"] - fw["mag_hat"]).max()))

    loss, g_y, g_m = eng.loss(y_hat, _t(y), mag_hat, _t(sbf), 2e-5 / 10)
    assert abs(loss.item() - ref_loss) < 1e-4 + 4 * abs(emu_loss - ref_loss)
    grads = [torch.full_like(p, float("nan")) for p in params]
    eng.backward(g_y, None, g_m, params, grads)
    for (name, _), gt in zip(O.param_order(d), grads):
        got, ref = gt.cpu().numpy(), ref_grads[name]
        assert np.isfinite(got).all(), name
        scale = np.abs(ref).max()
        emu_e = np.abs(emu_grads[name] - ref).max()
        e = np.abs(got - ref).max()
        # 3e-2 of the tensor's max-abs -- except where the emulation itself predicts more (denoise: the analysis gradients are
        # tiny and dominated by bins whose phase is ill-conditioned; emulation and CUDA path agree there, 8.6e-2 both)
        assert e <= max(GRAD_RTOL_TF32 * scale, 2 * emu_e) + 1e-12, (name, e / scale, emu_e / scale)
        assert e <= 6 * emu_e + 2e-3 * scale + 1e-12, (name, e / scale, emu_e / scale)

    # back to the default: bit-identical to the run before the switch
    eng.set_precision("fp32")
    y32b, _, mh32b, _ = eng.forward(_t(x), _t(knobs), params)
    assert torch.equal(y32, y32b) and torch.equal(mh32, mh32b)
    assert np.abs(y32b.cpu().numpy() - fw["y_hat"]).max() < 1e-5


def test_tf32_train_steps_track_the_fp32_trainer():
    """Three fused train steps in each mode from the same state: the loss trajectories agree to the reduced mode's
    error level and the parameters stay within Adam's step size of each other (lr * steps * 2)."""
    from signaltrain_b200.engine import Engine, Geometry
    d = O.model_dims(1, 4, 4)
    P = O.init_params(d, seed=218)
    rng = np.random.RandomState(3)
    B = 8
    t = np.arange(d.C) / 44100.0
    x = (0.5 * np.sin(2 * np.pi * rng.uniform(80, 2000, (B, 1)) * t) + 0.05 * rng.standard_normal((B, d.C))).astype(np.float32)
    y = np.tanh(1.5 * x[:, -d.L:]).astype(np.float32)
    knobs = rng.uniform(-0.5, 0.5, (B, d.K)).astype(np.float32)
    sbf = _t(O.scale_by_freq(d.F))
    lr, out = 1e-4, {}
    for mode in ("fp32", "tf32"):
        eng = Engine(Geometry(1, 4, 4), "cuda:0")
        eng.set_precision(mode)
        params = _dev_params(P, d)
        grads, m, v = ([torch.zeros_like(p) for p in params] for _ in range(3))
        losses = []
        for step in range(1, 4):
            losses.append(eng.train_step(_t(x), _t(y), _t(knobs), params, grads, m, v, sbf, 2e-6,
                                         eng.adam_hp(lr=lr, step=step, max_norm=1.0)).item())
        out[mode] = (losses, [p.cpu().numpy() for p in params])
    l32, l19 = np.array(out["fp32"][0]), np.array(out["tf32"][0])
    assert np.all(np.isfinite(l19))
    np.testing.assert_allclose(l19, l32, rtol=2e-2, atol=1e-4)
    for a, b in zip(out["fp32"][1], out["tf32"][1]):
        assert np.abs(a - b).max() <= 2 * 3 * lr * 1.01


def test_model_and_train_expose_the_mode():
    from signaltrain_b200 import nn_proc
    model = nn_proc.st_model(scale_factor=1, shrink_factor=4, num_knobs=4).to("cuda:0")
    model.set_precision("tf32")
    x = torch.randn(2, model.in_chunk_size, device="cuda:0") * 0.3
    k = torch.rand(2, 4, device="cuda:0") - 0.5
    y_tf32 = model.forward(x, k)[0]
    eng = next(iter(model.mpaec._engines.values()))
    assert eng.precision == "tf32"
    model.set_precision("fp32")
    y_fp32 = model.forward(x, k)[0]
    diff = (y_tf32 - y_fp32).abs().max().item()
    assert 1e-7 < diff < WAVE_TOL_TF32, diff
    with pytest.raises(ValueError):
        model.set_precision("fp8")
