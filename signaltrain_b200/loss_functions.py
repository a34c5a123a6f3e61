"""Host-side mirror of `signaltrain/loss_functions.py`: `logcosh` (:9-10), `mae` (:22-23), `calc_loss` (:26-43).
The reductions and their gradients are one CUDA kernel each (csrc/st_loss_opt.cu)."""
import torch

from .engine import Engine, Geometry

_engines = {}


def _engine_on(t):
    """An engine on the tensor's device: the model's if one lives there, else a small default-geometry handle kept per
    device.  The loss / MAE reductions take their sizes from the tensors (st_loss_shaped), not from the handle."""
    if not t.is_cuda:
        raise RuntimeError("signaltrain_b200.loss_functions: CUDA tensors only (no CPU fallback)")
    dev = t.device if t.device.index is not None else torch.device("cuda", torch.cuda.current_device())
    eng = Engine.any_on(dev)
    if eng is None:
        eng = _engines.get(dev.index)
        if eng is None:
            eng = _engines[dev.index] = Engine(Geometry(1, 4, 1), dev)
    return eng


class _CalcLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, y, mag_hat, sbf, l1_coef):
        eng = _engine_on(y_hat)
        loss, g_y, g_m = eng.loss_shaped(y_hat.contiguous(), y.contiguous(), mag_hat.contiguous(), sbf, l1_coef, want_grads=True)
        ctx.save_for_backward(g_y, g_m)
        return loss

    @staticmethod
    def backward(ctx, g):
        g_y, g_m = ctx.saved_tensors
        return g_y * g, None, g_m * g, None, None


def _freq_vector(scale_by_freq, mag_hat):
    """The reference passes exp(7 f/F) expanded over (B, OT, F) (train.py:115-117); the kernel wants the (F,) vector."""
    if scale_by_freq is None:
        return None
    s = scale_by_freq
    F = mag_hat.shape[-1]
    if s.dim() > 1:
        if any(st != 0 for st in s.stride()[:-1]) and s.shape[:-1].numel() > 1:
            first = s.reshape(-1, F)[0]
            if not torch.equal(s.reshape(-1, F), first.expand(s.reshape(-1, F).shape)):
                raise NotImplementedError("calc_loss: scale_by_freq must vary along the frequency axis only")
            s = first
        else:
            s = s.reshape(-1, F)[0] if s.is_contiguous() else s[(0,) * (s.dim() - 1)]
    return s.to(device=mag_hat.device, dtype=torch.float32).contiguous()


def calc_loss(y_hat, y_cuda, mag_hat, batch_size=20, scale_by_freq=None, l1_lambda=2e-5, reg_logcosh=False):
    """mean(log cosh(y - y_hat)) + l1 * mean|mag_hat (* scale_by_freq)|   (live branches :34 and :36)."""
    if reg_logcosh:
        raise NotImplementedError("reg_logcosh=True is not on the live path (train.py:120 never sets it)")
    sbf = _freq_vector(scale_by_freq, mag_hat)
    coef = l1_lambda if sbf is None else l1_lambda / 10
    return _CalcLoss.apply(y_hat, y_cuda.float(), mag_hat, sbf, coef)


def logcosh(y_hat, y):
    eng = _engine_on(y_hat)
    mh = torch.zeros((y_hat.shape[0], 1, 4), device=y_hat.device)
    loss, _, _ = eng.loss_shaped(y_hat.contiguous(), y.float().contiguous(), mh, None, 0.0, want_grads=False)
    return loss


def mae(x, x_hat):
    if not x.is_cuda:
        raise RuntimeError("signaltrain_b200.loss_functions: CUDA tensors only (no CPU fallback)")
    eng = _engine_on(x)
    return eng.mae(x.float().contiguous(), x_hat.float().contiguous())
