"""Checkpoint wire format of the path: `save_checkpoint` / `load_checkpoint` with the argument lists of
the reference's `signaltrain/misc.py:21,38`.  A checkpoint is one `torch.save`d dict -- the 40-tensor `state_dict` under the
reference's keys plus run metadata -- so files written here load in the reference's tools (`predict_long`, the demo) and vice
versa.  Unlike the reference (TODO at `train.py:229`) the optimizer state stored in the file is restored by `train()`."""
import collections
import os
import sys

import numpy as np
import torch

# run values assumed for old checkpoints that do not carry them (the reference's guesses, misc.py:50-57)
LEGACY_RUN_VALUES = {
    'sr': 44100, 'scale_factor': 1, 'shrink_factor': 4, 'in_chunk_size': 8192, 'out_chunk_size': 2048,
    'knob_names': ['thresh', 'ratio', 'attackTime', 'releaseTime'],
    'knob_ranges': np.array([[-30, 0], [1, 5], [1e-3, 4e-2], [1e-3, 4e-2]]),
}
MODEL_ATTRS = ('scale_factor', 'shrink_factor', 'in_chunk_size', 'out_chunk_size')
EFFECT_ATTRS = (('effect_name', 'name'), ('knob_names', 'knob_names'), ('knob_ranges', 'knob_ranges'))


def print_choochoo(version):
    print(f"signaltrain_b200 {version}  (B200-native train step; API of SignalTrain)\n")


def save_checkpoint(checkpointname, model, epoch, parallel, optimizer, effect, sr):
    core = model.module if parallel else model
    record = {field: getattr(effect, attr) for field, attr in EFFECT_ATTRS}
    record.update({name: getattr(core, name) for name in MODEL_ATTRS})
    record['sr'] = sr
    record['epoch'] = epoch + 1
    record['optimizer'] = optimizer.state_dict()
    # an OrderedDict like nn.Module.state_dict(), with CPU copies so that the file loads on any box
    record['state_dict'] = collections.OrderedDict((key, t.detach().cpu()) for key, t in core.state_dict().items())
    print(f'\nsaving model to {checkpointname}', end="")
    torch.save(record, checkpointname)


def load_checkpoint(checkpointname, fatal=False, device="cuda"):
    """-> (state_dict, run_values).  Both are empty when there is no such file (and `fatal` is off)."""
    if not os.path.isfile(checkpointname):
        if fatal:
            print("Error, no checkpoint found")
            sys.exit(1)
        return {}, {}
    print("\n***** Checkpoint file found. Loading weights.")
    record = torch.load(checkpointname, map_location=device, weights_only=False)
    run_values = dict(LEGACY_RUN_VALUES)
    run_values.update({key: value for key, value in record.items() if 'state_dict' not in key})
    return record['state_dict'], run_values
