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
