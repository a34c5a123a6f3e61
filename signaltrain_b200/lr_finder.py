"""Learning-rate range test: host-side mirror of the reference's `utils/lr_finder.py:18-55` (`lrfind`), the second caller
of the train step (forward, calc_loss, backward, clip_grad_norm_, optimizer.step through the mirrored module API, i.e.
the CUDA path).  The plot at the end of the reference's function is out of scope: the (lrs, losses) it would plot are
returned instead."""
import numpy as np


def lrfind(model, dataloader, optimizer, calc_loss, start=1e-6, stop=4e-3, num_lrs=150, to_screen=False, device="cuda:0"):
    lrs, losses = [], []
    lr_tries = np.logspace(np.log10(start), np.log10(stop), num_lrs)
    ind, count, repeat = 0, 0, 3
    for x, y, knobs in dataloader:
        count += 1
        if ind >= len(lr_tries):
            break
        lr_try = lr_tries[ind]
        if count % repeat == 0:              # repeat over this many data points per lr value
            ind += 1
        optimizer.param_groups[0]['lr'] = lr_try
        x_cuda, y_cuda, knobs_cuda = x.to(device), y.to(device), knobs.to(device)
        x_hat, mag, mag_hat = model.forward(x_cuda, knobs_cuda)
        loss = calc_loss(x_hat.float(), y_cuda.float(), mag.float())         # lr_finder.py:38 passes `mag`, no frequency weights
        lrs.append(lr_try)
        losses.append(loss.item())
        optimizer.zero_grad()
        loss.backward()
        model.clip_grad_norm_()
        optimizer.step()
    return lrs, losses
