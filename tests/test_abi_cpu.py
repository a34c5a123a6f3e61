"""CPU-side checks of the boundary: the shared library loads and exports every symbol the header declares,
the Python binding names the same set, and the product path refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "signaltrain_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", src)))


def test_header_binding_and_library_agree():
    from signaltrain_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    hdr = _header_symbols()
    assert len(hdr) >= 20
    assert hdr == _lib.exported_symbols()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in hdr:
        assert hasattr(lib, s), s
    lib.st_abi_version.restype = ctypes.c_int
    assert lib.st_abi_version() == 1


def test_sass_is_sm100a_only():
    from signaltrain_b200 import _lib
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    import signaltrain_b200 as st
    model = st.nn_proc.st_model(1, 4, 4)
    x = torch.zeros(2, model.in_chunk_size)
    k = torch.zeros(2, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.forward(x, k)
    with pytest.raises(RuntimeError, match="CUDA"):
        st.loss_functions.calc_loss(torch.zeros(2, 2048), torch.zeros(2, 2048), torch.zeros(2, 9, 513))


def test_model_surface_matches_reference_contract():
    """attribute names, state_dict keys and shapes."""
    import signaltrain_b200 as st
    from oracle import st_oracle as O
    for scale, K in ((1, 4), (2, 2), (1, 1)):
        m = st.nn_proc.st_model(scale_factor=scale, shrink_factor=4, num_knobs=K)
        d = O.model_dims(scale, 4, K)
        assert (m.in_chunk_size, m.out_chunk_size, m.num_knobs) == (d.C, d.L, K)
        assert (m.scale_factor, m.shrink_factor) == (scale, 4)
        sd = m.state_dict()
        assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, tuple(s)) for k, s in O.param_order(d)]
        assert m.mpaec.dft_analysis.conv_analysis_real.weight.shape == (1024, 1, 1024)
        assert hasattr(m, "clip_grad_norm_") and hasattr(m.mpaec, "clip_grad_norm_")


def test_sliding_window_mirror():
    """audio.py:23-49, the reference's own docstring example."""
    from signaltrain_b200.predict_long import sliding_window
    np.testing.assert_array_equal(sliding_window(np.arange(10), 5, overlap=2), [[0, 1, 2, 3, 4], [3, 4, 5, 6, 7], [6, 7, 8, 9, 0]])
    assert sliding_window(np.arange(8), 4, overlap=0).shape == (2, 4)


def test_learningrate_mirror_matches_reference_probes():
    """learningrate.py:14-52: the host LUT of the product package against values of the reference's own function
    (stored with the goldens by tests/golden/make_goldens.py)."""
    import signaltrain_b200 as st
    from tests.helpers import load_case
    g, _ = load_case("comp4c_c8192_k4_b3")
    lrs, moms = st.learningrate.get_1cycle_schedule(lr_max=1e-4, n_data_points=200000, epochs=1000, batch_size=200)
    assert len(lrs) == len(moms) == int(g["meta/lr_sched_len"])
    np.testing.assert_allclose(lrs[:8], g["meta/lr_sched_head"], rtol=1e-12)
    np.testing.assert_allclose(lrs[g["meta/lr_sched_probe_idx"]], g["meta/lr_sched_probe"], rtol=1e-12)
    assert moms.min() >= 0.85 - 1e-12 and moms.max() <= 0.95 + 1e-12 and np.isclose(moms[0], 0.95)
