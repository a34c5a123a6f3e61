"""Host-side mirror of `signaltrain/loss_functions.py`: `logcosh` (:9-10), `mae` (:22-23), `calc_loss` (:26-43).
The reductions and their gradients are one CUDA kernel each (csrc/st_loss_opt.cu)."""
import torch

from .engine import Engine, Geometry

_engines = {}


def _engine_like(y_hat, mag_hat=None):
    """A loss-only engine for tensors that did not come from a model on this device (geometry only matters
    through L and (OT, F), which are read off the tensor shapes)."""
    if not y_hat.is_cuda:
        raise RuntimeError("signaltrain_b200.loss_functions: CUDA tensors only (no CPU fallback)")
    dev = y_hat.device if y_hat.device.index is not None else torch.device("cuda", torch.cuda.current_device())
    found = Engine.find(dev, int(y_hat.shape[1]), *((int(mag_hat.shape[1]), int(mag_hat.shape[2])) if mag_hat is not None else ()))
    if found is not None:          # tensors produced by a model on this device: share its engine
        return found
    key = (y_hat.device.index, tuple(y_hat.shape[1:]), None if mag_hat is None else tuple(mag_hat.shape[1:]))
    eng = _engines.get(key)
    if eng is None:
        g = Geometry.__new__(Geometry)
        L = int(y_hat.shape[1])
        if mag_hat is not None:
            OT, F = int(mag_hat.shape[1]), int(mag_hat.shape[2])
        else:
            OT, F = 9, 513
        N = 2 * (F - 1)
        H = (L + N) // (OT - 1)
        if (OT - 1) * H - N != L:
            raise RuntimeError(f"calc_loss: shapes y_hat {tuple(y_hat.shape)} / mag_hat "
                               f"{None if mag_hat is None else tuple(mag_hat.shape)} do not describe a SignalTrain model")
        C = None
        for T in range(OT, 65):      # any chunk consistent with (N, H, T) will do for a loss-only handle
            c = (T - 1) * H - N
            if c >= L and c % 4 == 0 and (c + N) // H + 1 == T:
                C = c
                break
        if C is None:
            raise RuntimeError("calc_loss: cannot derive a consistent geometry for a loss-only engine")
        g.C, g.N, g.H, g.T, g.OT, g.L, g.F, g.K, g.R = C, N, H, T, OT, L, F, 1, 64
        g.intended_out_chunk = L
        eng = Engine(g, y_hat.device)
        _engines[key] = eng
    return eng


class _CalcLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, y, mag_hat, sbf, l1_coef):
        eng = _engine_like(y_hat, mag_hat)
        need = y_hat.requires_grad or mag_hat.requires_grad
        loss, g_y, g_m = eng.loss(y_hat.contiguous(), y.contiguous(), mag_hat.contiguous(), sbf, l1_coef, want_grads=True)
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
    eng = _engine_like(y_hat)
    mh = torch.zeros((y_hat.shape[0], eng.g.OT, eng.g.F), device=y_hat.device)
    loss, _, _ = eng.loss(y_hat.contiguous(), y.float().contiguous(), mh, None, 0.0, want_grads=False)
    return loss


def mae(x, x_hat):
    if not x.is_cuda:
        raise RuntimeError("signaltrain_b200.loss_functions: CUDA tensors only (no CPU fallback)")
    eng = next(iter(_engines.values()), None) or _engine_like(x if x.dim() == 2 else x.reshape(x.shape[0], -1))
    return eng.mae(x.float().contiguous(), x_hat.float().contiguous())
