"""DCT / MDCT front-end variant (reference: signaltrain/cls_fe_dct_bases.py).  CPU: the numpy oracle against goldens
minted from the reference's own layers.  GPU: the CUDA path (st_dct_analysis / st_dct_synthesis through the mirrored
Analysis / Synthesis / tied_transform classes) against the oracle and the goldens."""
import os

import numpy as np
import pytest
import torch

from oracle import st_oracle as O
from tests.helpers import perturbation

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["dct_ft256_w512_h256_c4096_b3", "dct_ft1024_w2048_h1024_c8192_b2"]


def _load(case):
    g = np.load(os.path.join(HERE, "golden", case + ".npz"))
    ft, w, hop = int(g["ft_size"]), int(g["w_size"]), int(g["hop"])
    core = O.dct_core_modulation(ft, w)
    Wa = core + perturbation((ft, 1, w), int(g["pert_seeds"][0]), 1e-3)[:, 0]
    Ws = core + perturbation((ft, 1, w), int(g["pert_seeds"][1]), 1e-3)[:, 0]
    return g, ft, w, hop, core, Wa.astype(np.float32), Ws.astype(np.float32)


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference(case):
    g, ft, w, hop, core, Wa, Ws = _load(case)
    np.testing.assert_allclose(core[g["core_rows"]], g["core_modulation_rows"], atol=1e-7, rtol=0)
    x_ft = O.dct_analysis_forward(g["x"], Wa, g["bias"], hop)
    assert x_ft.shape == g["x_ft"].shape
    np.testing.assert_allclose(x_ft, g["x_ft"], atol=3e-5, rtol=0)
    np.testing.assert_allclose(O.dct_synthesis_forward(g["x_ft"], Ws, hop), g["wave"], atol=3e-5, rtol=0)
    np.testing.assert_allclose(O.dct_synthesis_forward(g["x_ft"], Wa, hop), g["tied"], atol=3e-5, rtol=0)


def test_oracle_cosine_basis_reconstructs():
    """Property of the un-perturbed basis (the TDAC of an MDCT with a sine window): synthesis(analysis(x)) == x away from
    the chunk edges, with zero bias."""
    ft, w, hop, C = 64, 128, 64, 1024
    core = O.dct_core_modulation(ft, w)
    x = np.random.RandomState(0).standard_normal((2, C))
    y = O.dct_synthesis_forward(O.dct_analysis_forward(x, core, np.zeros(ft), hop), core, hop)[:, 0]
    np.testing.assert_allclose(y, x, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_matches_oracle_and_golden(case):
    import signaltrain_b200 as st
    g, ft, w, hop, core, Wa, Ws = _load(case)
    torch.manual_seed(218)
    ana = st.cls_fe_dct_bases.Analysis(ft_size=ft, w_size=w, hop_size=hop).cuda()
    syn = st.cls_fe_dct_bases.Synthesis(ft_size=ft, w_size=w, hop_size=hop).cuda()
    assert list(ana.state_dict().keys()) == ["conv_analysis.weight", "conv_analysis.bias"]
    assert list(syn.state_dict().keys()) == ["conv_synthesis.weight"]
    # same constructor order under the same seed as the golden script -> the reference's random Conv1d bias, bit for bit
    np.testing.assert_array_equal(ana.conv_analysis.bias.detach().cpu().numpy(), g["bias"])
    np.testing.assert_allclose(ana.conv_analysis.weight.detach().cpu().numpy()[:, 0], core, atol=1e-7, rtol=0)
    with torch.no_grad():
        ana.conv_analysis.weight.copy_(torch.from_numpy(Wa[:, None, :]))
        syn.conv_synthesis.weight.copy_(torch.from_numpy(Ws[:, None, :]))
    x_ft = ana.forward(g["x"])                                    # numpy in, like the reference's forward
    assert tuple(x_ft.shape) == g["x_ft"].shape
    ref = O.dct_analysis_forward(g["x"], Wa, g["bias"], hop)
    assert np.abs(x_ft.cpu().numpy() - ref).max() < 5e-6 + 3e-6 * np.abs(ref).max()
    np.testing.assert_allclose(x_ft.cpu().numpy(), g["x_ft"], atol=3e-5, rtol=0)
    wave = syn.forward(torch.from_numpy(g["x_ft"]).cuda())
    assert tuple(wave.shape) == g["wave"].shape
    assert np.abs(wave.cpu().numpy() - O.dct_synthesis_forward(g["x_ft"], Ws, hop)).max() < 1e-5     # the north star's waveform bar
    np.testing.assert_allclose(wave.cpu().numpy(), g["wave"], atol=3e-5, rtol=0)
    tied = st.cls_fe_dct_bases.tied_transform(ana, torch.from_numpy(g["x_ft"]).cuda(), hop)
    assert np.abs(tied.cpu().numpy() - O.dct_synthesis_forward(g["x_ft"], Wa, hop)).max() < 1e-5
    with pytest.raises(RuntimeError):
        syn.forward(torch.from_numpy(g["x_ft"]))                  # CPU tensor: no fallback


@pytest.mark.gpu
def test_cuda_round_trip_full_size():
    """Size-independent property at the benchmark batch (B=200, chunk 8192): the cosine basis reconstructs its input."""
    import signaltrain_b200 as st
    ana = st.cls_fe_dct_bases.Analysis().cuda()
    syn = st.cls_fe_dct_bases.Synthesis().cuda()
    with torch.no_grad():
        ana.conv_analysis.bias.zero_()
    x = torch.randn(200, 8192, device="cuda") * 0.3
    y = syn.forward(ana.forward(x))[:, 0]
    assert float((y - x).abs().max()) < 1e-5
