#!/usr/bin/env python3
"""Mint the wire-format manifest of a checkpoint written by the UNMODIFIED reference (read-only at /root/reference):
build its st_model + torch.optim.Adam as signaltrain/train.py:216-228 does, take one CPU train step so the optimizer carries
state, call its own misc.save_checkpoint (misc.py:21-35), reload the file and record its structure -- top-level fields and
types, state_dict keys / shapes / dtypes, optimizer.state_dict() layout, metadata values.  The 50 MB file itself is not kept.

Run here (CPU container) only:   python tests/golden/make_checkpoint_manifest.py
Writes tests/golden/checkpoint_manifest.json (checked by tests/test_checkpoint_format.py)."""
import json
import os
import sys
import tempfile
from unittest import mock

import numpy as np
import scipy.signal
import scipy.signal.windows
import torch

REF = os.environ.get("SIGNALTRAIN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
scipy.signal.hamming = scipy.signal.windows.hamming       # the four environment shims of SURVEY.md section 8(c)
scipy.signal.cosine = scipy.signal.windows.cosine
torch.has_cudnn = False
for name in ("librosa", "matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
    sys.modules.setdefault(name, mock.MagicMock())
sys.path.insert(0, REF)
import signaltrain as st  # noqa: E402  (the reference package)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers import checkpoint_manifest as manifest_of  # noqa: E402  (shared with tests/test_checkpoint_format.py)


def main():
    np.random.seed(218)
    torch.manual_seed(218)
    effect = st.audio.Compressor_4c()
    model = st.nn_proc.st_model(scale_factor=1, shrink_factor=4, num_knobs=len(effect.knob_names), sr=44100)
    optimizer = torch.optim.Adam(list(model.parameters()), lr=1e-4 / 15, weight_decay=0)        # train.py:228
    x = torch.randn(2, model.in_chunk_size) * 0.2
    knobs = torch.rand(2, 4) - 0.5
    y = torch.tanh(x[:, -model.out_chunk_size:])
    y_hat, mag, mag_hat = model.forward(x, knobs)
    loss = st.loss_functions.calc_loss(y_hat, y, mag_hat)
    optimizer.zero_grad()
    loss.backward()
    model.clip_grad_norm_()
    optimizer.step()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "modelcheckpoint.tar")
        st.misc.save_checkpoint(path, model, 4, False, optimizer, effect, 44100)
        size = os.path.getsize(path)
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    man = manifest_of(ckpt)
    man["file_bytes"] = size
    man["torch"] = torch.__version__
    with open(os.path.join(HERE, "checkpoint_manifest.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    print("wrote checkpoint_manifest.json:", len(man["state_dict"]), "tensors,", size, "bytes on disk")


if __name__ == "__main__":
    main()
