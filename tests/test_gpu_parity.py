"""GPU parity: the CUDA path (through the C ABI, via signaltrain_b200.engine) against the numpy oracle and the
golden vectors minted from the unmodified reference.  Tolerances are stated per check; the north-star bar is
1e-5 max-abs on output waveforms in fp32."""
import numpy as np
import pytest
import torch

from oracle import st_oracle as O
from tests.conftest import GOLDEN_CASES
from tests.helpers import assert_params_close_after_adam, dft_summary, initial_params, load_case

pytestmark = pytest.mark.gpu

WAVE_TOL = 1e-5          # BASELINE.json north_star: output waveforms within 1e-5 max-abs of the reference
SPEC_TOL = 5e-6          # spectra are O(1..30); fp32 contraction over 1024 taps: atol + SPEC_RTOL * |x|
SPEC_RTOL = 3e-6
GRAD_RTOL = 3e-4         # gradients: relative to the tensor's max-abs (fp32 sums over up to B*T*F terms)


def _engine(d, dev="cuda:0"):
    from signaltrain_b200.engine import Engine, Geometry
    g = Geometry.__new__(Geometry)
    g.C, g.N, g.H, g.T, g.OT, g.L, g.F, g.K, g.R = d.C, d.N, d.H, d.T, d.OT, d.L, d.F, d.K, d.R
    g.intended_out_chunk = d.L
    return Engine(g, dev)


def _dev_params(P, d, dev="cuda:0"):
    return [torch.from_numpy(np.ascontiguousarray(P[name])).to(dev) for name, _ in O.param_order(d)]


def _t(a, dev="cuda:0"):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def test_library_loaded_is_in_tree():
    from signaltrain_b200 import _lib
    lib = _lib.load()
    assert lib.st_abi_version() == 1
    assert "signaltrain_b200/lib/libsignaltrain_b200.so" in _lib.LIB_PATH


def test_init_frontend_matches_oracle():
    d = O.model_dims(1, 4, 4)
    eng = _engine(d)
    ws = [torch.empty((d.N, 1, d.N), device="cuda:0") for _ in range(4)]
    eng.init_frontend(ws)
    for w, ref in zip(ws, O.dft_init(d.N, d.H)):
        np.testing.assert_allclose(w.cpu().numpy()[:, 0], ref, atol=4e-9, rtol=0)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_forward_vs_oracle_and_golden(case):
    g, d = load_case(case)
    P = initial_params(g, d)
    eng = _engine(d)
    x, knobs = g["step0/x"], g["step0/knobs"]
    y_hat, mag, mag_hat, acts = eng.forward(_t(x), _t(knobs), _dev_params(P, d), return_acts=True)
    fw = O.forward(d, P, x, knobs, dtype=np.float64)
    c = lambda t: t.cpu().numpy()
    np.testing.assert_allclose(c(acts[0]), fw["re"], atol=SPEC_TOL, rtol=SPEC_RTOL)
    np.testing.assert_allclose(c(acts[1]), fw["im"], atol=SPEC_TOL, rtol=SPEC_RTOL)
    np.testing.assert_allclose(c(mag), fw["mag"], atol=SPEC_TOL, rtol=SPEC_RTOL)
    np.testing.assert_allclose(c(acts[2]), fw["mag"], atol=SPEC_TOL, rtol=SPEC_RTOL)
    ok = fw["mag"] > 1e-3                              # phase is ill-conditioned where the bin is empty
    dphi = np.abs(np.angle(np.exp(1j * (c(acts[3]) - fw["phs"]))))
    assert dphi[ok].max() < 2e-3
    # AE internals follow the reference's return_acts order (nn_proc.py:80-120)
    for base, cache in ((4, fw["mc"]), (14, fw["pc"])):
        a = cache["acts"]
        ref = [a[1], a[2], a[3], a[4][:, :, :16], a[4], a[5], a[6], a[7], a[8]]
        for i, r in enumerate(ref):
            tol = 2e-5 if base == 4 else 2e-3       # the phase AE also sees the ill-conditioned bins
            np.testing.assert_allclose(c(acts[base + i]), r, atol=tol, err_msg=f"act {base + i}")
    np.testing.assert_allclose(c(mag_hat), fw["mag_hat"], atol=2e-5)
    np.testing.assert_allclose(c(acts[24]), fw["mag_hat"], atol=2e-5)
    np.testing.assert_allclose(c(acts[26]), fw["an_re"], atol=2e-5)
    np.testing.assert_allclose(c(acts[27]), fw["an_im"], atol=2e-5)
    np.testing.assert_allclose(c(acts[28]), fw["x_fwdsyn"], atol=WAVE_TOL)
    np.testing.assert_allclose(c(acts[29]), fw["y_half"], atol=WAVE_TOL)
    np.testing.assert_allclose(c(y_hat), fw["y_hat"], atol=WAVE_TOL)
    np.testing.assert_allclose(c(y_hat), g["step0/y_hat"], atol=WAVE_TOL)       # the reference itself
    # production path (return_acts=False): tensor-core autoencoder kernels
    y2, mag2, mh2, _ = eng.forward(_t(x), _t(knobs), _dev_params(P, d))
    np.testing.assert_allclose(c(y2), fw["y_hat"], atol=WAVE_TOL)
    np.testing.assert_allclose(c(y2), g["step0/y_hat"], atol=WAVE_TOL)
    np.testing.assert_allclose(c(mag2), fw["mag"], atol=SPEC_TOL, rtol=SPEC_RTOL)
    np.testing.assert_allclose(c(mh2), fw["mag_hat"], atol=2e-5)
    np.testing.assert_allclose(eng.debug_read("phs_hat").reshape(fw["phs_hat"].shape)[fw["mag_hat"] > 1e-3],
                               fw["phs_hat"][fw["mag_hat"] > 1e-3], atol=2e-3)
    np.testing.assert_allclose(c(mag_hat), g["step0/mag_hat"], atol=2e-5)
    np.testing.assert_allclose(c(mag), g["step0/mag"], atol=SPEC_TOL, rtol=SPEC_RTOL)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_loss_and_backward_vs_oracle_and_golden(case):
    g, d = load_case(case)
    P = initial_params(g, d)
    eng = _engine(d)
    x, y, knobs = g["step0/x"], g["step0/y"].astype(np.float32), g["step0/knobs"]
    params = _dev_params(P, d)
    y_hat, mag, mag_hat, _ = eng.forward(_t(x), _t(knobs), params)
    sbf = O.scale_by_freq(d.F)
    loss, g_y, g_m = eng.loss(y_hat, _t(y), mag_hat, _t(sbf), 2e-5 / 10)
    ref_loss, ref_grads, fw = O.loss_and_grads(d, P, x, y, knobs, sbf)
    assert abs(loss.item() - ref_loss) < 2e-6
    assert abs(loss.item() - float(g["step0/loss"])) < 5e-6
    np.testing.assert_allclose(g_y.cpu().numpy(), -np.tanh(y - fw["y_hat"]) / fw["y_hat"].size, atol=1e-9)
    grads = [torch.full_like(p, float("nan")) for p in params]
    eng.backward(g_y, None, g_m, params, grads)
    for (name, _), gt in zip(O.param_order(d), grads):
        got = gt.cpu().numpy()
        ref = ref_grads[name].reshape(got.shape)
        assert np.isfinite(got).all(), name
        scale = max(np.abs(ref).max(), 1e-12)
        assert np.abs(got - ref).max() / scale < GRAD_RTOL, (name, np.abs(got - ref).max(), scale)
        if name in O.DFT_KEYS:                   # and against the reference's own autograd
            rows, sums = dft_summary(got, d.N)
            rr = g[f"step0/grad/{name}/rows"]
            assert np.abs(rows - rr).max() / max(np.abs(rr).max(), 1e-12) < 1e-3, name
        else:
            rr = g[f"step0/grad/{name}"]
            assert np.abs(got - rr).max() / max(np.abs(rr).max(), 1e-12) < 1e-3, name
    # clip (nn_proc.py:299-302)
    total = eng.clip_grad_norm(grads[:4], 1.0).item()
    ref_total = O.clip_grad_norm_(ref_grads)
    assert abs(total - ref_total) / ref_total < 2e-4
    for k, gt in zip(O.DFT_KEYS, grads[:4]):
        ref = ref_grads[k].reshape(gt.shape)
        assert np.abs(gt.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-12) < 5e-4


def test_external_grad_inputs_are_honoured():
    """g_mag (gradient w.r.t. the returned mag) and a non-loss g_mag_hat flow through st_backward."""
    g, d = load_case("comp4c_c8192_k4_b3")
    P = initial_params(g, d)
    eng = _engine(d)
    x, knobs = g["step0/x"], g["step0/knobs"]
    params = _dev_params(P, d)
    y_hat, mag, mag_hat, _ = eng.forward(_t(x), _t(knobs), params)
    rng = np.random.RandomState(7)
    gy = rng.standard_normal(y_hat.shape).astype(np.float32) * 1e-3
    gm = rng.standard_normal(mag.shape).astype(np.float32) * 1e-4
    gmh = rng.standard_normal(mag_hat.shape).astype(np.float32) * 1e-4
    grads = [torch.empty_like(p) for p in params]
    eng.backward(_t(gy), _t(gm), _t(gmh), params, grads)
    fw = O.forward(d, P, x, knobs, dtype=np.float64)
    ref = O.backward(d, fw, gy.astype(np.float64), gmh.astype(np.float64), gm.astype(np.float64))
    for (name, _), gt in zip(O.param_order(d), grads):
        r = ref[name].reshape(gt.shape)
        # random (incoherent) upstream gradients: heavy cancellation in the fp32 sums, hence the looser bound
        assert np.abs(gt.cpu().numpy() - r).max() / max(np.abs(r).max(), 1e-12) < 1e-3, name


def test_adam_step_vs_oracle():
    d = O.model_dims(1, 4, 4)
    eng = _engine(d)
    rng = np.random.RandomState(3)
    P = {n: (rng.standard_normal(s) * 0.1).astype(np.float32) for n, s in O.param_order(d)}
    G = {n: (rng.standard_normal(s) * 1e-3).astype(np.float32) for n, s in O.param_order(d)}
    names = [n for n, _ in O.param_order(d)]
    p = [_t(P[n]) for n in names]
    gr = [_t(G[n]) for n in names]
    m = [torch.zeros_like(t) for t in p]
    v = [torch.zeros_like(t) for t in p]
    state = {}
    Pd = {k: a.astype(np.float64) for k, a in P.items()}
    for step in (1, 2, 3):
        eng.adam_step(p, gr, m, v, eng.adam_hp(lr=3e-4, step=step))
        Pd = O.adam_step(Pd, {k: a.astype(np.float64) for k, a in G.items()}, state, 3e-4)
    for n, t in zip(names, p):
        np.testing.assert_allclose(t.cpu().numpy(), Pd[n], atol=2e-7, rtol=1e-6, err_msg=n)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_three_fused_train_steps(case):
    """st_train_step x3 (train.py:104-151 semantics incl. the one-step lr lag) vs the reference's parameters."""
    g, d = load_case(case)
    P = initial_params(g, d)
    eng = _engine(d)
    params = _dev_params(P, d)
    grads = [torch.zeros_like(p) for p in params]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    lr_sched, _ = O.get_1cycle_schedule(1e-4, 200000, 1000, 200)
    sbf = _t(O.scale_by_freq(d.F))
    tr = O.Trainer(d, P, lr_sched, dtype=np.float64)
    lr = float(lr_sched[0])
    for step in range(3):
        x, y, knobs = g[f"step{step}/x"], g[f"step{step}/y"], g[f"step{step}/knobs"]
        hp = eng.adam_hp(lr=lr, step=step + 1, max_norm=1.0)
        loss = eng.train_step(_t(x), _t(y), _t(knobs), params, grads, m, v, sbf, 2e-5 / 10, hp)
        lr = float(lr_sched[min(step, len(lr_sched) - 1)])
        ref_loss, _, _ = tr.step(x, y, knobs)
        assert abs(loss.item() - ref_loss) < 5e-6, (step, loss.item(), ref_loss)
        assert abs(loss.item() - float(g[f"step{step}/loss"])) < 1e-5
        if step in (0, 2):
            for (name, _), pt in zip(O.param_order(d), params):
                got = pt.cpu().numpy()
                if name in O.DFT_KEYS:
                    rows, _ = dft_summary(got, d.N)
                    assert_params_close_after_adam(rows, g[f"step{step}/params_after/{name}/rows"], name)
                    assert_params_close_after_adam(got, tr.P[name], name, frac=1e-4)
                else:
                    assert_params_close_after_adam(got, g[f"step{step}/params_after/{name}"], name, atol=7e-6, frac=0.02)


def test_module_api_path_matches_golden():
    """The reference's own call sequence (train.py:112-151) through the mirrored Python classes."""
    import signaltrain_b200 as st
    g, d = load_case("comp4c_c8192_k4_b3")
    torch.manual_seed(218)
    model = st.nn_proc.st_model(scale_factor=1, shrink_factor=4, num_knobs=4)
    assert model.in_chunk_size == d.C and model.out_chunk_size == d.L
    model.to("cuda:0")
    lr_sched, mom_sched = st.learningrate.get_1cycle_schedule(lr_max=1e-4, n_data_points=200000, epochs=1000, batch_size=200)
    opt = st.optim.Adam(model, lr=lr_sched[0], weight_decay=0)
    sbf = None
    for step in range(3):
        x, y, knobs = (torch.from_numpy(g[f"step{step}/{k}"]).cuda() for k in ("x", "y", "knobs"))
        y_hat, mag, mag_hat = model.forward(x, knobs)
        if sbf is None:
            expfac = 7. / mag_hat.size()[-1]
            sbf = torch.exp(expfac * torch.arange(0., mag_hat.size()[-1])).expand_as(mag_hat).float()
        loss = st.loss_functions.calc_loss(y_hat.float(), y.float(), mag_hat.float(), scale_by_freq=sbf)
        assert abs(loss.item() - float(g[f"step{step}/loss"])) < 1e-5
        np.testing.assert_allclose(y_hat.detach().cpu().numpy(), g[f"step{step}/y_hat"], atol=2e-5)
        opt.zero_grad()
        loss.backward()
        model.clip_grad_norm_()
        opt.step()
        opt.param_groups[0]['lr'] = lr_sched[min(step, len(lr_sched) - 1)]
        opt.param_groups[0]['momentum'] = mom_sched[min(step, len(mom_sched) - 1)]
    sd = model.state_dict()
    assert list(sd.keys()) == [n for n, _ in O.param_order(d)]
    for name, _ in O.param_order(d):
        got = sd[name].cpu().numpy()
        if name in O.DFT_KEYS:
            rows, _ = dft_summary(got, d.N)
            assert_params_close_after_adam(rows, g[f"step2/params_after/{name}/rows"], name)
        else:
            assert_params_close_after_adam(got, g[f"step2/params_after/{name}"], name, atol=7e-6, frac=0.02)


def test_full_size_batch_properties():
    """BASELINE configs[1] size (B=200, C=8192, K=4): size-independent properties + oracle waveform parity."""
    d = O.model_dims(1, 4, 4)
    eng = _engine(d)
    P = O.init_params(d, seed=11)
    rng = np.random.RandomState(5)
    B = 200
    t = np.arange(d.C) / 44100.0
    x = (0.4 * np.sin(2 * np.pi * rng.uniform(50, 4000, (B, 1)) * t + rng.uniform(0, 6.28, (B, 1)))
         + 0.05 * rng.standard_normal((B, d.C))).astype(np.float32)
    knobs = (rng.beta(0.8, 0.8, (B, d.K)) - 0.5).astype(np.float32)
    params = _dev_params(P, d)
    xd, kd = _t(x), _t(knobs)
    y1, mag1, mh1, _ = eng.forward(xd, kd, params)
    y2, _, _, _ = eng.forward(xd, kd, params)
    assert torch.equal(y1, y2)                                   # run-to-run determinism
    # windows are independent: two half batches reproduce the full batch bit for bit
    ya, _, _, _ = eng.forward(xd[:100].contiguous(), kd[:100].contiguous(), params)
    yb, _, _, _ = eng.forward(xd[100:].contiguous(), kd[100:].contiguous(), params)
    assert torch.equal(torch.cat([ya, yb]), y1)
    fw = O.forward(d, P, x, knobs, dtype=np.float32, keep=False)
    np.testing.assert_allclose(y1.cpu().numpy(), fw["y_hat"], atol=WAVE_TOL)
    np.testing.assert_allclose(mh1.cpu().numpy(), fw["mag_hat"], atol=5e-5)
    # gradient of a mean loss: full batch == average of the two halves (up to fp32 summation order)
    y = np.tanh(1.3 * x[:, -d.L:]).astype(np.float32)
    sbf = _t(O.scale_by_freq(d.F))

    def grads_of(sl):
        yh, _, mh, _ = eng.forward(xd[sl].contiguous(), kd[sl].contiguous(), params)
        _, gy, gm = eng.loss(yh, _t(y[sl]), mh, sbf, 2e-6)
        gs = [torch.empty_like(p) for p in params]
        eng.backward(gy, None, gm, params, gs)
        return gs
    gf, ga, gb = grads_of(slice(0, 200)), grads_of(slice(0, 100)), grads_of(slice(100, 200))
    for i, (f, a, b) in enumerate(zip(gf, ga, gb)):
        avg = 0.5 * (a + b)
        assert (f - avg).abs().max().item() <= 2e-4 * max(f.abs().max().item(), 1e-12), i
    # dead analysis rows (bins >= F are sliced off, cls_fe_dft.py:55-56) get exactly zero gradient;
    # synthesis gradients are exactly Hermitian
    assert gf[0][d.F:].abs().max().item() == 0.0 and gf[1][d.F:].abs().max().item() == 0.0
    sr, si = gf[2][:, 0], gf[3][:, 0]
    assert torch.equal(sr[1:d.F - 1], torch.flip(sr[d.F:], [0]))
    assert torch.equal(si[1:d.F - 1], -torch.flip(si[d.F:], [0]))


def test_standalone_analysis_synthesis_roundtrip():
    """Analysis.forward / Synthesis.forward as stand-alone classes (cls_fe_dft.py:50-58, 102-115): oracle parity, and
    the STFT pair at its Fourier/Griffin-Lim initialisation reconstructs the waveform (8e-7)."""
    import signaltrain_b200 as st
    d = O.model_dims(1, 4, 4)
    rng = np.random.RandomState(2)
    B = 5
    x = (0.5 * np.sin(2 * np.pi * 440 * np.arange(d.C) / 44100.0) + 0.1 * rng.standard_normal((B, d.C))).astype(np.float32)
    ana = st.cls_fe_dft.Analysis(d.N, d.H).cuda()
    syn = st.cls_fe_dft.Synthesis(d.N, d.H).cuda()
    re, im = ana.forward(torch.from_numpy(x).cuda())
    Wr, Wi, Sr, Si = O.dft_init(d.N, d.H)
    ref_re, ref_im, _ = O.analysis_forward(d, Wr.astype(np.float64), Wi.astype(np.float64), x.astype(np.float64))
    np.testing.assert_allclose(re.cpu().numpy(), ref_re, atol=SPEC_TOL, rtol=SPEC_RTOL)
    np.testing.assert_allclose(im.cpu().numpy(), ref_im, atol=SPEC_TOL, rtol=SPEC_RTOL)
    wave = syn.forward(re, im)                              # all T frames -> (T-1)*H - N samples back
    assert wave.shape == (B, (d.T - 1) * d.H - d.N)
    n = wave.shape[1]
    np.testing.assert_allclose(wave.cpu().numpy(), x[:, :n], atol=5e-6)


def test_autoencoder_backward_implementations_agree(monkeypatch):
    """The tcgen05 / TMEM backward with in-kernel recompute (production), the mma.sync backward from saved records and the
    recomputing SIMT backward are three independent implementations of the same gradients: they must agree on a ragged batch
    (37 windows: the last tile of every kernel is partial) and on a large one (512, BASELINE configs[2] shape)."""
    d = O.model_dims(1, 4, 4)
    P = O.init_params(d, seed=3)
    rng = np.random.RandomState(9)
    for B in (37, 512):
        t = np.arange(d.C) / 44100.0
        x = (0.4 * np.sin(2 * np.pi * rng.uniform(50, 4000, (B, 1)) * t) + 0.05 * rng.standard_normal((B, d.C))).astype(np.float32)
        knobs = (rng.beta(0.8, 0.8, (B, d.K)) - 0.5).astype(np.float32)
        y = np.tanh(1.3 * x[:, -d.L:]).astype(np.float32)
        results = []
        for env in ({}, {"ST_DISABLE_TMEM_AE": "1"}, {"ST_DISABLE_TMEM_AE": "1", "ST_DISABLE_MMA_BACKWARD": "1"}):
            for k in ("ST_DISABLE_TMEM_AE", "ST_DISABLE_MMA_BACKWARD"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            eng = _engine(d)                                  # the switches are read when the handle is created
            params = _dev_params(P, d)
            yh, _, mh, _ = eng.forward(_t(x), _t(knobs), params)
            _, gy, gm = eng.loss(yh, _t(y), mh, _t(O.scale_by_freq(d.F)), 2e-6)
            gs = [torch.zeros_like(p) for p in params]
            eng.backward(gy, None, gm, params, gs)
            torch.cuda.synchronize()
            results.append([g.clone() for g in gs])
            if B == 512 and len(results) == 2:
                break                                         # the SIMT kernel at B=512 adds nothing the mma one does not
        ref = results[0]
        for other in results[1:]:
            for i, (a, b) in enumerate(zip(ref, other)):
                assert (a - b).abs().max().item() <= 3e-4 * max(a.abs().max().item(), 1e-12), (B, i)


def test_backward_in_two_calls_is_bit_identical():
    """st_backward_begin + st_backward_finish (the data-parallel split: the synthesis gradients are final after `begin`)
    produce exactly the gradients of st_backward; after `begin` the synthesis pair already holds its final values."""
    d = O.model_dims(1, 4, 4)
    eng = _engine(d)
    P = O.init_params(d, seed=4)
    rng = np.random.RandomState(2)
    B = 9
    x = (0.3 * rng.standard_normal((B, d.C))).astype(np.float32)
    knobs = (rng.beta(0.8, 0.8, (B, d.K)) - 0.5).astype(np.float32)
    y = np.tanh(x[:, -d.L:]).astype(np.float32)
    params = _dev_params(P, d)
    sbf = _t(O.scale_by_freq(d.F))

    def run(split):
        yh, _, mh, _ = eng.forward(_t(x), _t(knobs), params)
        _, gy, gm = eng.loss(yh, _t(y), mh, sbf, 2e-6)
        gs = [torch.full_like(p, float("nan")) for p in params]
        if not split:
            eng.backward(gy, None, gm, params, gs)
            return gs, None
        eng.backward(gy, None, gm, params, gs, part="begin")
        torch.cuda.synchronize()
        early = [gs[2].clone(), gs[3].clone()]
        eng.backward(gy, None, gm, params, gs, part="finish")
        return gs, early
    whole, _ = run(False)
    parts, early = run(True)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(whole, parts)):
        assert torch.equal(a, b), i
    assert torch.equal(early[0], whole[2]) and torch.equal(early[1], whole[3])
    with pytest.raises(RuntimeError):
        eng.backward(_t(y), None, None, params, parts, part="finish")          # no begin before it


@pytest.mark.parametrize("case", ["comp4c_c8192_k4_b3", "comp2k_c16384_k2_b2"])
def test_fused_step_tail_matches_the_piecewise_entry_points(case):
    """st_train_step runs a fused forward tail (overlap-add + residual + loss + loss gradients + padded 2*dL/dy_hat in one
    kernel) and a fused DFT-gradient finalisation + L1 norm; st_forward / st_loss / st_backward / st_adam_step run the separate
    kernels.  Same arithmetic per element: all 40 gradients must be bit-identical, the loss and the clip coefficient differ
    only by summation order, so parameters after the update agree to one Adam step of fp32 noise."""
    g, d = load_case(case)
    P = initial_params(g, d)
    x, y, knobs = _t(g["step0/x"]), _t(g["step0/y"].astype(np.float32)), _t(g["step0/knobs"])
    sbf = _t(O.scale_by_freq(d.F))
    eng = _engine(d)
    out = []
    for fused in (True, "grad_step", False):
        params = _dev_params(P, d)
        grads = [torch.full_like(p, float("nan")) for p in params]
        m = [torch.zeros_like(p) for p in params]
        v = [torch.zeros_like(p) for p in params]
        hp = eng.adam_hp(lr=1e-4, step=1, max_norm=1.0)
        if fused is True:
            loss = eng.train_step(x, y, knobs, params, grads, m, v, sbf, 2e-6, hp)
        elif fused == "grad_step":       # the data-parallel route: st_grad_step, (allreduce), st_adam_step
            loss = eng.grad_step(x, y, knobs, params, grads, sbf, 2e-6)
            eng.adam_step(params, grads, m, v, hp)
        else:
            y_hat, _, mag_hat, _ = eng.forward(x, knobs, params)
            loss, g_y, g_m = eng.loss(y_hat, y, mag_hat, sbf, 2e-6)
            eng.backward(g_y, None, g_m, params, grads)
            eng.adam_step(params, grads, m, v, hp)
        out.append((loss.item(), [t.clone() for t in grads], [t.clone() for t in params]))
    (lf, gf, pf), (lg, gg, pg), (lu, gu, pu) = out
    assert lg == lf                                              # same loss kernel
    for i, (a, b) in enumerate(zip(gg, gu)):
        assert torch.equal(a, b), ("grad_step", i)
    for i, (a, b) in enumerate(zip(pg, pu)):
        assert torch.equal(a, b), ("grad_step params", i)        # same norm kernel as the piecewise route: bit-identical update
    assert abs(lf - lu) < 2e-7 * max(1.0, abs(lu)), (lf, lu)
    for i, (a, b) in enumerate(zip(gf, gu)):
        assert torch.equal(a, b), (i, (a - b).abs().max().item())
    for i, (a, b) in enumerate(zip(pf, pu)):
        assert (a - b).abs().max().item() <= (8e-9 if i < 4 else 0.0), (i, (a - b).abs().max().item())   # DFT: <= 2 ulp of 0.034


@pytest.mark.parametrize("B", [12, 37, 200])
def test_whole_step_vs_oracle_at_batches_that_use_pair_tiles(B):
    """The golden cases hold 2-3 windows (one 128-row tile: the 1-CTA GEMM kernel serves the forward there).  At B = 12 / 37 the
    forward GEMMs run on cta_group::2 tiles (324 / 999 frame rows = 3 / 8 tiles, i.e. a half-empty last pair / exact pairs), and
    the fused step tail sees a ragged batch; B = 200 is the bench size (BASELINE configs[1]: every SM busy, 5-6 autoencoder tiles per
    CTA in the forward and 10-11 in the backward).  Forward, loss and all 40 gradients (st_grad_step) against the float64 oracle."""
    d = O.model_dims(1, 4, 4)
    P = O.init_params(d, seed=218)
    rng = np.random.RandomState(B)
    t = np.arange(d.C) / 44100.0
    x = (0.4 * np.sin(2 * np.pi * rng.uniform(60, 3000, (B, 1)) * t + rng.uniform(0, 6.28, (B, 1)))
         + 0.05 * rng.standard_normal((B, d.C))).astype(np.float32)
    y = np.tanh(1.5 * x[:, -d.L:]).astype(np.float32)
    knobs = rng.uniform(-0.5, 0.5, (B, d.K)).astype(np.float32)
    sbf = O.scale_by_freq(d.F)
    eng = _engine(d)
    params = _dev_params(P, d)
    y_hat, _, mag_hat, _ = eng.forward(_t(x), _t(knobs), params)
    ref_loss, ref_grads, fw = O.loss_and_grads(d, P, x, y, knobs, sbf)
    np.testing.assert_allclose(y_hat.cpu().numpy(), fw["y_hat"], atol=WAVE_TOL)
    np.testing.assert_allclose(mag_hat.cpu().numpy(), fw["mag_hat"], atol=2e-5)
    grads = [torch.full_like(p, float("nan")) for p in params]
    loss = eng.grad_step(_t(x), _t(y), _t(knobs), params, grads, _t(sbf), 2e-5 / 10)
    assert abs(loss.item() - ref_loss) < 2e-6
    for (name, _), gt in zip(O.param_order(d), grads):
        got, ref = gt.cpu().numpy(), ref_grads[name]
        assert np.isfinite(got).all(), name
        assert np.abs(got - ref).max() <= GRAD_RTOL * np.abs(ref).max() + 1e-12, (name, np.abs(got - ref).max() / np.abs(ref).max())


def test_packed_exchange_payload_roundtrip():
    """st_pack_grads / st_unpack_grads (the data-parallel exchange payload: 8.5 of 16.8 MB): the packed buffer equals the
    host-side statement of the layout (parallel.pack_payload), and unpacking restores all 40 gradient tensors BIT-exactly --
    analysis rows >= F are exactly zero and the synthesis gradients exactly Hermitian after a real backward."""
    from signaltrain_b200 import parallel
    d = O.model_dims(1, 4, 4)
    eng = _engine(d)
    P = O.init_params(d, seed=5)
    rng = np.random.RandomState(3)
    B = 7
    x = (0.3 * rng.standard_normal((B, d.C))).astype(np.float32)
    knobs = (rng.beta(0.8, 0.8, (B, d.K)) - 0.5).astype(np.float32)
    y = np.tanh(x[:, -d.L:]).astype(np.float32)
    params = _dev_params(P, d)
    grads = [torch.full_like(p, float("nan")) for p in params]
    eng.grad_step(_t(x), _t(y), _t(knobs), params, grads, _t(O.scale_by_freq(d.F)), 2e-6)
    packed = torch.empty(eng.packed_grad_floats(), device="cuda")
    eng.pack_grads(grads, packed)
    ref = parallel.pack_payload(grads, d.F)
    n_dft = 4 * d.F * d.N
    assert torch.equal(packed[:n_dft], ref[:n_dft])
    o_lib, o_ref = n_dft, n_dft                                   # the library aligns every autoencoder tensor to 4 floats
    for g_ in grads[4:]:
        assert torch.equal(packed[o_lib:o_lib + g_.numel()], ref[o_ref:o_ref + g_.numel()])
        o_lib += (g_.numel() + 3) // 4 * 4
        o_ref += g_.numel()
    assert o_lib == packed.numel()
    restored = [torch.full_like(g_, float("nan")) for g_ in grads]
    for r_, g_ in zip(restored[:2], grads[:2]):
        r_.view(d.N, -1)[d.F:] = 0.0                              # dead analysis rows are not part of the payload: left untouched
    eng.unpack_grads(packed, restored)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(grads, restored)):
        assert torch.equal(a, b), i


def test_packed_grad_step_unpack_clip_and_clipped_adam_match_the_piecewise_calls():
    """The three calls of the data-parallel step (st_grad_step_packed -> [allreduce] -> st_unpack_clip -> st_adam_step_clipped)
    against st_grad_step + st_adam_step with the L1 clip: the restored gradients are BIT-identical (same split-K and partial
    sums, only the destination differs), the norm agrees to summation order, and the parameters after the update to 1e-7 --
    with a grad_scale != 1, as after an allreduce over two ranks."""
    d = O.model_dims(1, 4, 4)
    eng = _engine(d)
    P = O.init_params(d, seed=5)
    rng = np.random.RandomState(11)
    B = 9
    x = (0.3 * rng.standard_normal((B, d.C))).astype(np.float32)
    knobs = (rng.beta(0.8, 0.8, (B, d.K)) - 0.5).astype(np.float32)
    y = np.tanh(x[:, -d.L:]).astype(np.float32)
    sbf = _t(O.scale_by_freq(d.F))
    scale = 0.5

    def fresh():
        p = _dev_params(P, d)
        return p, [torch.zeros_like(t) for t in p], [torch.zeros_like(t) for t in p], [torch.zeros_like(t) for t in p]

    pa, ga, ma, va = fresh()
    loss_a = eng.grad_step(_t(x), _t(y), _t(knobs), pa, ga, sbf, 2e-6)
    ga_keep = [g_.clone() for g_ in ga]
    eng.adam_step(pa, ga, ma, va, eng.adam_hp(1e-3, 1, grad_scale=scale, max_norm=1.0))

    pb, gb, mb, vb = fresh()
    packed = torch.full((eng.packed_grad_floats(),), float("nan"), device="cuda")
    loss_b = eng.grad_step_packed(_t(x), _t(y), _t(knobs), pb, packed, sbf, 2e-6)
    total = torch.empty((), device="cuda")
    eng.unpack_clip(packed, gb, scale, 1.0, total)
    eng.adam_step_clipped(pb, gb, mb, vb, eng.adam_hp(1e-3, 1, grad_scale=scale, max_norm=1.0))
    torch.cuda.synchronize()
    assert loss_a.item() == loss_b.item()
    for i, (a, b) in enumerate(zip(ga_keep, gb)):
        assert torch.equal(a, b), i
    ref_norm = scale * sum(float(g_.double().abs().sum()) for g_ in ga_keep[:4])
    assert abs(total.item() - ref_norm) <= 1e-5 * ref_norm
    for i, (a, b) in enumerate(zip(pa, pb)):
        assert (a - b).abs().max().item() <= 1e-7 * max(1.0, a.abs().max().item()), i
    for i, (a, b) in enumerate(zip(ma + va, mb + vb)):
        assert (a - b).abs().max().item() <= 1e-6 * max(1e-12, a.abs().max().item()), i
