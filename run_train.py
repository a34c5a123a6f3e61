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

    parser = argparse.ArgumentParser(description="Trains neural network to reproduce input-output transformations.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--apex', help="precision, in the reference's apex vocabulary: O0 = fp32-faithful (3xTF32 tensor cores); O1/O2/O3 = single-pass TF32 products", default="O0")
    parser.add_argument('-b', '--batch', type=int, help="batch size", default=200)
    parser.add_argument('--checkpoint', help='Name of model checkpoint .tar file', default="modelcheckpoint.tar")
    parser.add_argument('-c', '--compand', help='Turn on to use companded/decompanded audio', action='store_true')
    parser.add_argument('--effect', help='Name of effect to use', default="comp_4c")
    parser.add_argument('--epochs', type=int, help='Number of epochs to run', default=1000)
    parser.add_argument('--lrmax', type=float, help="max learning rate", default=1e-4)
    parser.add_argument('-n', '--num', type=int, help='Number of "data points" (audio clips) per epoch', default=200000)
    parser.add_argument('--path', help='Directory to pull input (and maybe target) data from', default=None)
    parser.add_argument('--sr', type=int, help='Sampling rate', default=44100)
    parser.add_argument('--scale', type=float, help='Scale factor (of input size & whole model)', default=1.0)
    parser.add_argument('--shrink', type=int, help='Shink output chunk relative to input by this divisor', default=4)
    parser.add_argument('-t', '--target', help="type of target: chunk or stream", default="stream")
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
