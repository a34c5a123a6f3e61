"""CPU: the oracle's model of the reduced-precision mode (oracle.operand_rounding) -- rounding is TF32 round-to-nearest
ties-away as cvt.rna.tf32.f32, it is off by default, and the error it predicts for the path sits where
tests/test_gpu_precision.py's stated bounds assume (a few 1e-4 on output waveforms, below 1e-2 of each gradient's max)."""
import numpy as np

from oracle import st_oracle as O
from tests.helpers import initial_params, load_case


def test_round_tf32_bit_patterns():
    a = np.array([1.0, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -12, 1.0 + 2.0 ** -10, -1.0 - 2.0 ** -11, 0.0, 3.0e-39], dtype=np.float32)
    r = O.round_tf32(a)
    assert r.dtype == np.float32
    np.testing.assert_array_equal(r[:6], np.array([1.0, 1.0 + 2.0 ** -10, 1.0, 1.0 + 2.0 ** -10, -1.0 - 2.0 ** -10, 0.0], np.float32))
    assert np.all((r.view(np.uint32) & np.uint32(0x1FFF)) == 0)
    x = np.random.RandomState(0).standard_normal(10000).astype(np.float32)
    rel = np.abs(O.round_tf32(x) - x) / np.abs(x)
    assert rel.max() <= 2.0 ** -11 * (1 + 1e-6)
    assert O.round_tf32(x.astype(np.float64)).dtype == np.float64


def test_rounding_is_off_by_default_and_restored():
    assert O._OPERAND_ROUNDING is None
    with O.operand_rounding("tf32"):
        assert O._OPERAND_ROUNDING == "tf32"
        with O.operand_rounding(None):
            assert O._OPERAND_ROUNDING is None
        assert O._OPERAND_ROUNDING == "tf32"
    assert O._OPERAND_ROUNDING is None


def test_predicted_error_level_of_the_reduced_mode():
    g, d = load_case("comp4c_c8192_k4_b3")
    P = initial_params(g, d)
    x, y, knobs = g["step0/x"], g["step0/y"].astype(np.float32), g["step0/knobs"]
    sbf = O.scale_by_freq(d.F)
    l0, g0, f0 = O.loss_and_grads(d, P, x, y, knobs, sbf)
    with O.operand_rounding("tf32"):
        l1, g1, f1 = O.loss_and_grads(d, P, x, y, knobs, sbf)
    l2, _, f2 = O.loss_and_grads(d, P, x, y, knobs, sbf)
    assert l2 == l0 and np.array_equal(f2["y_hat"], f0["y_hat"])            # leaving the context restores exact arithmetic
    e = np.abs(f1["y_hat"] - f0["y_hat"]).max()
    assert 1e-6 < e < 1e-3, e
    assert abs(l1 - l0) < 1e-4
    for k in g0:
        assert np.abs(g1[k] - g0[k]).max() <= 1.5e-2 * np.abs(g0[k]).max() + 1e-12, k
