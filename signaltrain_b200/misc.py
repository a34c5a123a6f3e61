"""Checkpoint I/O: mirror of `signaltrain/misc.py` `save_checkpoint` (:21-35) and `load_checkpoint` (:38-66).
Same dictionary fields and state_dict keys, so files interchange with the reference's tools."""
import os
import sys

import numpy as np
import torch


def print_choochoo(version):
    print(f"signaltrain_b200 {version}  (B200-native train step; API of SignalTrain)\n")


def save_checkpoint(checkpointname, model, epoch, parallel, optimizer, effect, sr):
    print(f'\nsaving model to {checkpointname}', end="")
    m = model.module if parallel else model
    state = {'epoch': epoch + 1,
             'state_dict': {k: v.detach().cpu() for k, v in m.state_dict().items()},
             'optimizer': optimizer.state_dict(),
             'effect_name': effect.name, 'knob_names': effect.knob_names, 'knob_ranges': effect.knob_ranges,
             'scale_factor': m.scale_factor, 'shrink_factor': m.shrink_factor,
             'in_chunk_size': m.in_chunk_size, 'out_chunk_size': m.out_chunk_size, 'sr': sr}
    torch.save(state, checkpointname)


def load_checkpoint(checkpointname, fatal=False, device="cuda"):
    """Returns (state_dict, run_values); both empty when the file does not exist.  Missing run values are
    filled with the reference's defaults (misc.py:50-57)."""
    state_dict, rv = {}, {}
    if os.path.isfile(checkpointname):
        print("\n***** Checkpoint file found. Loading weights.")
        checkpoint = torch.load(checkpointname, map_location=device, weights_only=False)
        state_dict = checkpoint['state_dict']
        rv = {'sr': 44100, 'scale_factor': 1, 'shrink_factor': 4, 'in_chunk_size': 8192, 'out_chunk_size': 2048,
              'knob_names': ['thresh', 'ratio', 'attackTime', 'releaseTime'],
              'knob_ranges': np.array([[-30, 0], [1, 5], [1e-3, 4e-2], [1e-3, 4e-2]])}
        for key, value in checkpoint.items():
            if 'state_dict' not in key:
                rv[key] = value
    elif fatal:
        print("Error, no checkpoint found")
        sys.exit(1)
    return state_dict, rv
