"""BASELINE configs[0] plumbing (run_train.py --effect comp_4c --epochs 1 --batch 4 --num 64) through the mirrored
train() / train_loop() / eval_status_save(): side files, checkpoint wire format, resume."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_one_epoch_writes_reference_artifacts(tmp_path, monkeypatch):
    import signaltrain_b200 as st
    from oracle import st_oracle as O
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(218)
    np.random.seed(218)
    model = st.train.train(effect=st.data.Compressor_4c(), epochs=1, n_data_points=64, batch_size=4,
                           device=torch.device("cuda:0"), lr_max=1e-4)
    assert os.path.exists("vl_avg_out.dat") and os.path.exists("val_err_mae.dat") and os.path.exists("modelcheckpoint.tar")
    ep, val = open("vl_avg_out.dat").read().split()
    assert ep == "1" and 0.0 < float(val) < 1.0
    ck = torch.load("modelcheckpoint.tar", map_location="cpu", weights_only=False)
    d = O.model_dims(1, 4, 4)
    assert list(ck["state_dict"].keys()) == [n for n, _ in O.param_order(d)]          # reference wire format (misc.py:28-34)
    for k in ("epoch", "optimizer", "effect_name", "knob_names", "knob_ranges", "scale_factor", "shrink_factor",
              "in_chunk_size", "out_chunk_size", "sr"):
        assert k in ck
    assert ck["in_chunk_size"] == 8192 and ck["out_chunk_size"] == 2048 and ck["epoch"] == 1
    # the weights moved, stayed finite, and a resumed run starts from them
    w0 = O.dft_init(d.N, d.H)[0]
    w1 = ck["state_dict"][O.DFT_KEYS[0]].numpy()[:, 0]
    assert np.isfinite(w1).all() and 1e-6 < np.abs(w1 - w0).max() < 1e-2
    model2 = st.train.train(effect=st.data.Compressor_4c(), epochs=1, n_data_points=16, batch_size=4,
                            device=torch.device("cuda:0"), lr_max=1e-4)
    assert np.abs(model2.state_dict()[O.DFT_KEYS[0]].cpu().numpy()[:, 0] - w1).max() < 1e-3


def test_validation_forward_does_not_need_grad():
    import signaltrain_b200 as st
    torch.manual_seed(218)
    model = st.nn_proc.st_model(1, 4, 4).cuda().eval()
    x = torch.randn(3, model.in_chunk_size, device="cuda") * 0.1
    k = torch.zeros(3, 4, device="cuda")
    with torch.no_grad():
        y1, mag, mh = model.forward(x, k)
    y2, _, _ = model.forward(x, k)
    assert torch.equal(y1, y2) and y1.shape == (3, model.out_chunk_size)
    assert not y1.requires_grad and y2.requires_grad


def test_host_batch_pipeline_matches_step_by_step():
    """run_host_batches (H2D of batch i+1 overlapped with step i, loss read one step late) is the same arithmetic as
    calling step() on device copies of the same batches: bit-identical losses and parameters."""
    import signaltrain_b200 as st
    from signaltrain_b200.train import FusedTrainer
    lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)
    pool = st.data.make_pool(5 * 6, 8192, 2048, st.data.Compressor_4c(), 44100, seed=3)
    host = [tuple(torch.from_numpy(a[i * 6:(i + 1) * 6]).pin_memory() for a in pool) for i in range(5)]
    results = []
    for mode in range(2):
        torch.manual_seed(218)
        model = st.nn_proc.st_model(1, 4, 4).cuda()
        tr = FusedTrainer(model, lr)
        if mode == 0:
            losses = [float(tr.step(*(t.cuda() for t in b)).item()) for b in host]
        else:
            seen = []
            losses = tr.run_host_batches(host, on_loss=lambda i, v: seen.append((i, v)))
            assert [i for i, _ in seen] == list(range(5)) and [v for _, v in seen] == losses
        torch.cuda.synchronize()
        results.append((losses, [p.detach().clone() for p in model.parameters()]))
    assert results[0][0] == results[1][0]
    for a, b in zip(results[0][1], results[1][1]):
        assert torch.equal(a, b)
    assert tr.run_host_batches([]) == []


def test_predict_long_matches_oracle_windows():
    """utils/predict_long.py:30-79: a long signal through overlapping windows, forward only; output bookkeeping as the reference."""
    import signaltrain_b200 as st
    from oracle import st_oracle as O
    from signaltrain_b200.predict_long import predict_long, sliding_window
    torch.manual_seed(218)
    model = st.nn_proc.st_model(1, 4, 4).cuda()
    d = O.model_dims(1, 4, 4)
    n = 8192 + 5 * 2048 + 777                                     # not a whole number of hops: exercises the zero padding
    rng = np.random.RandomState(5)
    sig = (0.4 * np.sin(2 * np.pi * 440 * np.arange(n) / 44100) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    knobs = np.array([0.1, -0.3, 0.25, 0.4], dtype=np.float32)
    y = predict_long(sig, knobs, model, model.in_chunk_size, model.out_chunk_size, device="cuda:0", batch_size=4)
    assert y.dtype == np.float64
    win = sliding_window(sig, 8192, overlap=8192 - 2048)
    P = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    fw = O.forward(d, P, np.ascontiguousarray(win), np.tile(knobs, (win.shape[0], 1)), dtype=np.float64)
    ref = fw["y_hat"].reshape(-1)
    unique = 8192 + (win.shape[0] - 1) * 2048
    ref = ref[:-(unique - n)] if unique > n else ref
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() < 1e-5
    assert model.training                                          # restored


def test_lrfind_runs_the_train_step_through_the_module_api():
    """utils/lr_finder.py:18-55: geometric lr sweep, three batches per value; the loss moves and stays finite."""
    import signaltrain_b200 as st
    from signaltrain_b200.lr_finder import lrfind
    torch.manual_seed(218)
    model = st.nn_proc.st_model(1, 4, 4).cuda()
    opt = st.optim.Adam(model, lr=1e-6)
    pool = st.data.make_pool(4 * 9, 8192, 2048, st.data.Compressor_4c(), 44100, seed=11)
    batches = [tuple(torch.from_numpy(a[i * 4:(i + 1) * 4]) for a in pool) for i in range(9)]
    lrs, losses = lrfind(model, batches, opt, st.loss_functions.calc_loss, start=1e-6, stop=1e-3, num_lrs=3)
    assert len(lrs) == len(losses) == 9 and lrs[0] == lrs[2] and lrs[3] > lrs[2] and np.isclose(lrs[-1], 1e-3)
    assert np.isfinite(losses).all() and len(set(losses)) > 1


def test_resume_restores_the_adam_step_count():
    """N + M steps in one run == N steps, checkpoint (model + optimizer state_dict), fresh objects, load, M more steps:
    bias correction continues from step N (optim.Adam.load_state_dict), so parameters agree bit for bit."""
    import signaltrain_b200 as st
    from signaltrain_b200.train import FusedTrainer
    lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)
    pool = st.data.make_pool(6 * 5, 8192, 2048, st.data.Compressor_4c(), 44100, seed=21)
    batches = [tuple(torch.from_numpy(a[i * 5:(i + 1) * 5]).cuda() for a in pool) for i in range(6)]
    N, M = 4, 2

    def fresh():
        torch.manual_seed(218)
        model = st.nn_proc.st_model(1, 4, 4).cuda()
        return model, st.optim.Adam(model, lr=float(lr[0]))

    model_a, opt_a = fresh()
    tr_a = FusedTrainer(model_a, lr, optimizer=opt_a)
    for b in batches[:N + M]:
        tr_a.step(*b)
    model_b, opt_b = fresh()
    tr_b = FusedTrainer(model_b, lr, optimizer=opt_b)
    for b in batches[:N]:
        tr_b.step(*b)
    tr_b.sync_optimizer_state()
    sd_model = {k: v.clone() for k, v in model_b.state_dict().items()}
    sd_opt = opt_b.state_dict()
    assert int(sd_opt["state"][0]["step"]) == N
    model_c, opt_c = fresh()
    model_c.load_state_dict(sd_model)
    opt_c.load_state_dict(sd_opt)
    assert opt_c._step == N
    tr_c = FusedTrainer(model_c, lr, optimizer=opt_c)
    tr_c.iter_count, tr_c.lr = tr_b.iter_count, tr_b.lr            # the schedule position is the caller's to restore
    opt_c.param_groups[0]["lr"] = tr_b.lr
    for b in batches[N:N + M]:
        tr_c.step(*b)
    torch.cuda.synchronize()
    for pa, pc in zip(model_a.parameters(), model_c.parameters()):
        assert torch.equal(pa, pc)


def test_eval_forward_between_steps_keeps_the_fast_path():
    """A no_grad validation forward (eval_status_save, predict_long) flips the engine to inference mode; the next train
    step must still run the tensor-core autoencoder kernels (no SIMT fallback) with the same number of launches."""
    import signaltrain_b200 as st
    from signaltrain_b200.train import FusedTrainer
    lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)
    torch.manual_seed(218)
    model = st.nn_proc.st_model(1, 4, 4).cuda()
    tr = FusedTrainer(model, lr)
    pool = st.data.make_pool(8, 8192, 2048, st.data.Compressor_4c(), 44100, seed=5)
    x, y, k = (torch.from_numpy(a).cuda() for a in pool)
    tr.step(x, y, k)
    eng = tr.eng
    n0 = eng.launch_count()
    tr.step(x, y, k)
    per_step = eng.launch_count() - n0
    with torch.no_grad():
        model.eval()
        model.forward(x, k)
        model.train()
    n1, f1 = eng.launch_count(), eng.fallback_count()
    tr.step(x, y, k)
    assert eng.launch_count() - n1 == per_step and eng.fallback_count() == f1 == 0


def test_lrfind_at_scale_factor_2():
    """utils/lr_finder.py:38 passes the INPUT magnitude (B, T, F) as mag_hat; at chunk 16384 (T = 46) that is not a geometry a
    handle could be built from -- calc_loss takes its sizes from the tensors (st_loss_shaped)."""
    import signaltrain_b200 as st
    from signaltrain_b200.lr_finder import lrfind
    torch.manual_seed(218)
    model = st.nn_proc.st_model(2, 4, 2).cuda()
    opt = st.optim.Adam(model, lr=1e-6)
    pool = st.data.make_pool(2 * 6, model.in_chunk_size, model.out_chunk_size, st.data.Compressor_2knob(), 44100, seed=13)
    batches = [tuple(torch.from_numpy(a[i * 2:(i + 1) * 2]) for a in pool) for i in range(6)]
    lrs, losses = lrfind(model, batches, opt, st.loss_functions.calc_loss, start=1e-6, stop=1e-4, num_lrs=2)
    assert len(losses) == 6 and np.isfinite(losses).all()


def test_library_calls_leave_the_current_device_alone():
    """Every C-ABI entry point runs on the handle's device and restores the caller's current device (DeviceGuard)."""
    import signaltrain_b200 as st
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.manual_seed(218)
    model = st.nn_proc.st_model(1, 4, 4).to("cuda:1")
    torch.cuda.set_device(0)
    x = torch.randn(2, model.in_chunk_size, device="cuda:1") * 0.1
    k = torch.zeros(2, 4, device="cuda:1")
    with torch.no_grad():
        model.forward(x, k)
    assert torch.cuda.current_device() == 0
    assert float(st.loss_functions.mae(x, x)) == 0.0 and torch.cuda.current_device() == 0


def test_train_from_prerecorded_files(tmp_path, monkeypatch):
    """train(datapath=...) (train.py:240-246): Train/ and Val/ file pairs preloaded into HBM, windows cropped on the device;
    with target_type != "stream" the 4-knob compressor is re-run on every chunk on the device (datasets.py:241-242)."""
    import signaltrain_b200 as st
    from tests.test_data_step import _write_pairs
    monkeypatch.chdir(tmp_path)
    rng = np.random.RandomState(1)
    for sub in ("Train", "Val"):
        _write_pairs(str(tmp_path / "data" / sub), 3, 12000, rng)
    torch.manual_seed(218)
    np.random.seed(218)
    model = st.train.train(effect=st.data.Compressor_4c(), epochs=1, n_data_points=40, batch_size=4, device=torch.device("cuda:0"),
                           datapath=str(tmp_path / "data"), lr_max=1e-4)
    ep, val = open("vl_avg_out.dat").read().split()
    assert ep == "1" and 0.0 < float(val) < 1.0
    assert all(torch.isfinite(p).all() for p in model.parameters())
    os.remove("modelcheckpoint.tar")
    model = st.train.train(effect=st.data.Compressor_4c(), epochs=1, n_data_points=40, batch_size=4, device=torch.device("cuda:0"),
                           datapath=str(tmp_path / "data"), target_type="chunk", lr_max=1e-4)
    assert all(torch.isfinite(p).all() for p in model.parameters())
    with pytest.raises(NotImplementedError):
        st.train.train(effect=st.data.Denoise(), epochs=1, n_data_points=40, batch_size=4, device=torch.device("cuda:0"),
                       datapath=str(tmp_path / "data"), target_type="chunk", in_checkpointname="none.tar")


def test_train_step_cuda_graph_replay_is_bit_identical(monkeypatch):
    """st_train_step captures itself into a CUDA graph on a capturable stream (second call with the same tensor tables) and
    replays it afterwards with this step's x / y / knobs / loss pointers and Adam scalars patched in: six steps on batches that
    live at different addresses give bit-identical parameters and losses to the plain launches (ST_CUDA_GRAPH=0), the legacy
    default stream is served without a graph."""
    import signaltrain_b200 as st
    from signaltrain_b200 import data
    from signaltrain_b200.train import FusedTrainer
    B, steps = 6, 6
    lr, _ = st.learningrate.get_1cycle_schedule(1e-4, 200000, 1000, 200)

    def run(graph, own_stream):
        monkeypatch.setenv("ST_CUDA_GRAPH", "1" if graph else "0")
        torch.manual_seed(218)
        model = st.nn_proc.st_model(1, 4, 4).cuda()
        tr = FusedTrainer(model, lr)
        x, y, k = (torch.from_numpy(a).cuda() for a in data.make_pool(B * steps, model.in_chunk_size, model.out_chunk_size,
                                                                      data.Compressor_4c(), seed=5))
        torch.cuda.synchronize()
        s = torch.cuda.Stream() if own_stream else torch.cuda.current_stream()
        losses = []
        with torch.cuda.stream(s):
            for i in range(steps):
                sl = slice(i * B, (i + 1) * B)
                losses.append(tr.step(x[sl], y[sl], k[sl]).clone())
        torch.cuda.synchronize()
        return [float(l) for l in losses], [p.detach().clone() for p in tr.params], tr.eng.graph_replays(), tr.eng.launch_count()

    l0, p0, r0, n0 = run(False, True)
    l1, p1, r1, n1 = run(True, True)
    l2, p2, r2, n2 = run(True, False)
    assert r0 == 0 and r2 == 0 and r1 == steps - 1          # step 1 plain, step 2 captured + launched, steps 3.. replayed
    assert n0 == n1 == n2                                    # the launch counter counts the kernels inside a replay
    assert l0 == l1 == l2
    for a, b, c in zip(p0, p1, p2):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_packed_grad_step_cuda_graph_replay_is_bit_identical(monkeypatch):
    """The data-parallel rank step (st_grad_step_packed) is captured and replayed the same way: payloads bit-identical to the
    plain launches on batches at different addresses."""
    import signaltrain_b200 as st
    from signaltrain_b200 import data
    B, steps = 5, 5

    def run(graph):
        monkeypatch.setenv("ST_CUDA_GRAPH", "1" if graph else "0")
        torch.manual_seed(218)
        model = st.nn_proc.st_model(1, 4, 4).cuda()
        x, y, k = (torch.from_numpy(a).cuda() for a in data.make_pool(B * steps, model.in_chunk_size, model.out_chunk_size,
                                                                      data.Compressor_4c(), seed=9))
        eng = model.mpaec._engine_for(x)
        params = [p.detach() for p in model.ordered_parameters()]
        sbf = torch.exp(torch.tensor(7.0 / eng.g.F) * torch.arange(0., eng.g.F)).float().cuda()
        packed = torch.empty(eng.packed_grad_floats(), device="cuda")
        out, losses = [], []
        torch.cuda.synchronize()
        with torch.cuda.stream(torch.cuda.Stream()):
            for i in range(steps):
                sl = slice(i * B, (i + 1) * B)
                losses.append(eng.grad_step_packed(x[sl], y[sl], k[sl], params, packed, sbf, 2e-6).clone())
                out.append(packed.clone())
        torch.cuda.synchronize()
        return out, [float(l) for l in losses], eng.graph_replays()

    a, la, ra = run(False)
    b, lb, rb = run(True)
    assert ra == 0 and rb == steps - 1 and la == lb
    for u, v in zip(a, b):
        assert torch.equal(u, v)
