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
