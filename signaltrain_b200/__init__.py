"""signaltrain_b200: B200-native train-step hot path of SignalTrain behind the reference's Python API.

Module names follow the reference package (`cls_fe_dft`, `cls_fe_dct_bases`, `nn_proc`, `loss_functions`, `learningrate`,
`train`, `misc`) so `import signaltrain_b200 as st` reads like `import signaltrain as st`."""
from . import cls_fe_dct_bases, cls_fe_dft, data, device_data, learningrate, loss_functions, misc, nn_proc, optim, parallel, predict_long, train  # noqa: F401

__version__ = "0.1.0"
