#!/usr/bin/env python3
"""Trainer entry point with the reference's command line (run_train.py:32-47 of drscotthawley/signaltrain) on the
B200 train step.  Same flags and defaults; the effect table maps to this repo's synthetic stand-ins
(signaltrain_b200/data.py) unless the reference package `signaltrain` is importable, in which case its own Effect
classes and datasets are used for the data layer (they are out of this repo's scope) and only the model / loss /
optimizer / loop come from here."""
import argparse
import glob
import sys

import numpy as np
import torch

import signaltrain_b200 as st

__version__ = st.__version__


def pick_effect(name, path):
    if name in st.data.EFFECTS:
        return st.data.EFFECTS[name]()
    try:                                   # the reference's data layer, if installed
        import signaltrain as ref
        table = {"files": lambda: ref.audio.FileEffect(path), "comp": ref.audio.Compressor, "comp_t": ref.audio.Comp_Just_Thresh,
                 "comp_large": ref.audio.Compressor_4c_Large, "comp_one": ref.audio.Compressor_4c_OneSetting,
                 "lowpass": ref.audio.LowPass}
        if name in table:
            return table[name]()
    except ImportError:
        pass
    print(f"Effect option '{name}' is not available (built in: {sorted(st.data.EFFECTS)}; the rest need the reference's audio.py)")
    sys.exit(1)


if __name__ == "__main__":
    np.random.seed(218)
    torch.manual_seed(218)
    if not torch.cuda.is_available():
        print("run_train.py: no CUDA device.  The B200 train step has no CPU fallback; run the reference's run_train.py on CPU.")
        sys.exit(2)
    device = torch.device("cuda:0")
    torch.cuda.manual_seed(218)

    # the reference's command line (run_train.py:32-47): same flags, short forms and defaults; help text is this repo's
    FLAGS = [
        (('--apex',), dict(default="O0", help="precision in apex vocabulary: O0 = fp32-faithful (3xTF32 tensor cores), "
                                              "O1/O2/O3 = single-pass TF32 products (fp32 storage, loss, optimiser)")),
        (('-b', '--batch'), dict(type=int, default=200, help="windows per step")),
        (('--checkpoint',), dict(default="modelcheckpoint.tar", help="checkpoint file to resume from / write to")),
        (('-c', '--compand'), dict(action='store_true', help="accepted for compatibility (companding is a data-layer option)")),
        (('--effect',), dict(default="comp_4c", help=f"audio effect to learn: {', '.join(sorted(st.data.EFFECTS))} built in")),
        (('--epochs',), dict(type=int, default=1000, help="passes over --num windows")),
        (('--lrmax',), dict(type=float, default=1e-4, help="peak of the 1-cycle learning-rate schedule")),
        (('-n', '--num'), dict(type=int, default=200000, help="windows per epoch")),
        (('--path',), dict(default=None, help="dataset directory (Train/ and Val/ with input_*/target_* files); none = synthetic")),
        (('--sr',), dict(type=int, default=44100, help="sample rate in Hz")),
        (('--scale',), dict(type=float, default=1.0, help="input chunk = 8192 x scale")),
        (('--shrink',), dict(type=int, default=4, help="output chunk = input chunk / shrink")),
        (('-t', '--target',), dict(default="stream", help="target alignment: chunk or stream")),
    ]
    parser = argparse.ArgumentParser(description="Train the SignalTrain model on a B200 (signaltrain_b200 train step).",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for names, spec in FLAGS:
        parser.add_argument(*names, **spec)
    args = parser.parse_args()
    print("Command line: ", " ".join(sys.argv[:]))

    effect = pick_effect(args.effect, args.path)
    if args.target not in ["chunk", "stream"]:
        print(f"Error, invalid target type: {args.target}")
        sys.exit(1)
    if args.path is not None and not glob.glob(args.path + "/Train/input*"):
        print(f"No input files under {args.path}/Train: synthesising data on the fly")
        args.path = None

    st.misc.print_choochoo(__version__)
    print("Running with args =", args)
    model = st.train.train(epochs=args.epochs, n_data_points=args.num, batch_size=args.batch, device=device, sr=args.sr,
                           effect=effect, datapath=args.path, scale_factor=args.scale, shrink_factor=args.shrink,
                           apex_opt=args.apex, target_type=args.target, lr_max=args.lrmax,
                           in_checkpointname=args.checkpoint, compand=args.compand)
    print("run_train.py: Execution completed.")
