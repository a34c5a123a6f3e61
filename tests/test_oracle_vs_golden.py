"""Pins oracle/st_oracle.py to outputs of the unmodified reference (tests/golden/*.npz, minted by
tests/golden/make_goldens.py).  CPU only."""
import numpy as np
import pytest

from oracle import st_oracle as O
from tests.conftest import GOLDEN_CASES
from tests.helpers import dft_summary, initial_params, load_case

ACT_TOL = 2e-5      # fp32 reference vs fp64 oracle, activations O(1..30)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_init_matches_reference(case):
    g, d = load_case(case)
    if "meta/perturb_seed" in g:
        pytest.skip("perturbed front-end: init rows are not the DFT init")
    for k, w in zip(O.DFT_KEYS, O.dft_init(d.N, d.H)):
        rows, sums = dft_summary(w, d.N)
        np.testing.assert_allclose(rows, g[f"init/{k}/rows"], atol=2e-7, rtol=0)
        np.testing.assert_allclose(sums, g[f"init/{k}/sums"], rtol=1e-5)
    names = [n for n, _ in O.param_order(d)]
    assert len(names) == 40


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_forward_acts(case):
    g, d = load_case(case)
    P = initial_params(g, d)
    fw = O.forward(d, P, g["step0/x"], g["step0/knobs"], dtype=np.float64)
    A = lambda n: g["step0/acts/" + n]
    np.testing.assert_allclose(fw["re"], A("x_real"), atol=ACT_TOL)
    np.testing.assert_allclose(fw["im"], A("x_imag"), atol=ACT_TOL)
    np.testing.assert_allclose(fw["mag"], A("mag"), atol=ACT_TOL)
    # atan2 is ill-conditioned where mag ~ 0 (all-padding frames): compare phase where it is defined
    ok = A("mag") > 1e-3
    dphi = np.abs(np.angle(np.exp(1j * (fw["phs"] - A("phs")))))
    assert dphi[ok].max() < 1e-3
    np.testing.assert_allclose(fw["mag_hat"], A("mag_hat"), atol=ACT_TOL)
    np.testing.assert_allclose(fw["an_re"], A("an_real"), atol=ACT_TOL)
    np.testing.assert_allclose(fw["an_im"], A("an_imag"), atol=ACT_TOL)
    np.testing.assert_allclose(fw["x_fwdsyn"], A("x_fwdsyn"), atol=ACT_TOL)
    np.testing.assert_allclose(fw["y_half"], A("y_hat_half"), atol=ACT_TOL)
    np.testing.assert_allclose(fw["y_hat"], g["step0/y_hat"], atol=1e-5)
    np.testing.assert_allclose(fw["mag"], g["step0/mag"], atol=ACT_TOL)
    # AE internals: b=0, every 16th bin.  acts[0] is the input, acts[i+1] the i-th ELU output;
    # reference list = 4 enc outputs, catted, 4 outputs, final out
    for tag, cache in (("m", fw["mc"]), ("p", fw["pc"])):
        acts = cache["acts"]
        ref_order = [acts[1], acts[2], acts[3], None, acts[4], acts[5], acts[6], acts[7], acts[8]]
        for i, a in enumerate(ref_order):
            if a is None:
                continue
            tol = ACT_TOL if tag == "m" else 5e-4     # phase AE sees the ill-conditioned bins too
            np.testing.assert_allclose(a[0, ::16], A(f"{tag}_act{i}"), atol=tol)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_loss_and_grads(case):
    g, d = load_case(case)
    P = initial_params(g, d)
    sbf = O.scale_by_freq(d.F)
    loss, grads, fw = O.loss_and_grads(d, P, g["step0/x"], g["step0/y"].astype(np.float32), g["step0/knobs"], sbf)
    assert abs(loss - float(g["step0/loss"])) < 2e-6
    assert abs(O.logcosh(fw["y_hat"], g["step0/y"].astype(np.float32)) - float(g["step0/logcosh"])) < 2e-6
    assert abs(O.mae(g["step0/y"].astype(np.float32), fw["y_hat"]) - float(g["step0/mae"])) < 2e-6
    for name, shape in O.param_order(d):
        if name in O.DFT_KEYS:
            rows, sums = dft_summary(grads[name], d.N)
            ref_rows = g[f"step0/grad/{name}/rows"]
            scale = max(np.abs(ref_rows).max(), 1e-12)
            assert np.abs(rows - ref_rows).max() / scale < 2e-4, name
            np.testing.assert_allclose(sums[0], g[f"step0/grad/{name}/sums"][0], rtol=2e-4)
        else:
            ref = g[f"step0/grad/{name}"]
            scale = max(np.abs(ref).max(), 1e-12)
            assert np.abs(grads[name] - ref).max() / scale < 5e-4, name
    total = O.clip_grad_norm_(grads)
    ref_total = sum(float(g[f"step0/grad/{k}/sums"][0]) for k in O.DFT_KEYS)
    assert abs(total - ref_total) / ref_total < 2e-4
    for k in O.DFT_KEYS:
        rows, _ = dft_summary(grads[k], d.N)
        ref_rows = g[f"step0/grad_clipped/{k}/rows"]
        assert np.abs(rows - ref_rows).max() / max(np.abs(ref_rows).max(), 1e-12) < 1e-3   # fp32 reference noise


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_three_train_steps(case):
    g, d = load_case(case)
    P = initial_params(g, d)
    lr_sched, _ = O.get_1cycle_schedule(1e-4, 200000, 1000, 200)
    assert len(lr_sched) == int(g["meta/lr_sched_len"])
    np.testing.assert_allclose(lr_sched[:8], g["meta/lr_sched_head"], rtol=1e-12)
    np.testing.assert_allclose(lr_sched[g["meta/lr_sched_probe_idx"]], g["meta/lr_sched_probe"], rtol=1e-12)
    tr = O.Trainer(d, P, lr_sched, dtype=np.float64)
    for step in range(3):
        loss, _, fw = tr.step(g[f"step{step}/x"], g[f"step{step}/y"], g[f"step{step}/knobs"])
        assert abs(loss - float(g[f"step{step}/loss"])) < 5e-6, (step, loss)
        np.testing.assert_allclose(fw["y_hat"], g[f"step{step}/y_hat"], atol=2e-5)
        if step in (0, 2):
            for name, _ in O.param_order(d):
                if name in O.DFT_KEYS:
                    rows, sums = dft_summary(tr.P[name], d.N)
                    np.testing.assert_allclose(rows, g[f"step{step}/params_after/{name}/rows"], atol=3e-6)
                else:
                    # Adam's first steps move every weight by ~lr regardless of gradient size, so a
                    # weight whose gradient is at fp32-noise level can differ in sign of the update
                    np.testing.assert_allclose(tr.P[name], g[f"step{step}/params_after/{name}"], atol=3 * 7e-6)
