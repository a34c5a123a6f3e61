"""CPU: checkpoint wire format.  A file written by signaltrain_b200.misc.save_checkpoint has the
structure of one written by the unmodified reference (tests/golden/checkpoint_manifest.json, minted by
tests/golden/make_checkpoint_manifest.py from the reference's own save_checkpoint after one of its train steps): same top-level
fields and value types, same 40 state_dict keys / shapes / dtypes in the same order, an optimizer entry torch.optim.Adam can load,
same metadata values -- so each side's tools read the other's files.  No GPU needed: nothing here runs the path."""
import json
import os

import numpy as np
import torch

from tests.helpers import GOLDEN_DIR, checkpoint_manifest


def _our_checkpoint(tmp_path):
    import signaltrain_b200 as st
    torch.manual_seed(218)
    effect = st.data.Compressor_4c()
    model = st.nn_proc.st_model(scale_factor=1, shrink_factor=4, num_knobs=len(effect.knob_names), sr=44100)
    optimizer = st.optim.Adam(model, lr=1e-4 / 15, weight_decay=0)
    optimizer._state_lists()                                  # what the first step creates: step / exp_avg / exp_avg_sq per tensor
    path = os.path.join(tmp_path, "modelcheckpoint.tar")
    st.misc.save_checkpoint(path, model, 4, False, optimizer, effect, 44100)
    return path, model, optimizer


def test_written_file_has_the_reference_structure(tmp_path):
    ref = json.load(open(os.path.join(GOLDEN_DIR, "checkpoint_manifest.json")))
    path, _, _ = _our_checkpoint(str(tmp_path))
    ours = json.loads(json.dumps(checkpoint_manifest(torch.load(path, map_location="cpu", weights_only=False))))
    assert ours["state_dict"] == ref["state_dict"]                      # keys, order, shapes, dtypes
    assert set(ours["top_level"]) == set(ref["top_level"])
    for field, desc in ref["top_level"].items():
        assert ours["top_level"][field]["type"] == desc["type"], field
        if "value" in desc:                                              # effect name, knob names / ranges, sizes, sr, epoch
            np.testing.assert_equal(ours["top_level"][field]["value"], desc["value"])
    ro, oo = ref["optimizer"], ours["optimizer"]
    assert oo["keys"] == ro["keys"] and oo["n_param_groups"] == ro["n_param_groups"] == 1
    assert oo["param_group_keys"] == ro["param_group_keys"]
    assert oo["params"] == ro["params"] == list(range(40))
    assert oo["n_state"] == ro["n_state"] == 40 and oo["state_shapes"] == ro["state_shapes"]
    assert oo["state_entry"] == ro["state_entry"]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize"):
        assert oo["param_group_values"][k] == ro["param_group_values"][k], k
    assert abs(os.path.getsize(path) - ref["file_bytes"]) < 64 * 1024    # same payload (40 tensors + 80 Adam moments)


def test_files_interchange_with_torch_adam_and_load_checkpoint(tmp_path):
    import signaltrain_b200 as st
    path, model, _ = _our_checkpoint(str(tmp_path))
    state_dict, rv = st.misc.load_checkpoint(path, device="cpu")
    assert rv["scale_factor"] == 1 and rv["shrink_factor"] == 4 and rv["sr"] == 44100 and rv["epoch"] == 5
    assert rv["effect_name"] == "Compressor_4c" and list(rv["knob_names"]) == ["threshold", "ratio", "attackTime", "releaseTime"]
    clone = st.nn_proc.st_model(scale_factor=rv["scale_factor"], shrink_factor=rv["shrink_factor"], num_knobs=len(rv["knob_names"]))
    clone.load_state_dict(state_dict)                                    # strict: every key present, nothing extra
    for (ka, a), (kb, b) in zip(model.state_dict().items(), clone.state_dict().items()):
        assert ka == kb and torch.equal(a, b)
    # the optimizer entry loads into the class the reference constructs (train.py:228) ...
    stock = torch.optim.Adam(list(clone.parameters()), lr=1.0, weight_decay=0)
    stock.load_state_dict(rv["optimizer"])
    assert stock.param_groups[0]["lr"] == 1e-4 / 15 and len(stock.state) == 40
    # ... and into this repo's (optimizer-state restore, which the reference leaves as a TODO at train.py:229)
    ours = st.optim.Adam(clone, lr=1.0)
    ours.load_state_dict(rv["optimizer"])
    assert ours.param_groups[0]["lr"] == 1e-4 / 15 and len(ours.state) == 40
    # a missing file: empty results, or exit when fatal (misc.py:63-65)
    assert st.misc.load_checkpoint(os.path.join(str(tmp_path), "nope.tar")) == ({}, {})
