"""Learning-rate range test: `lrfind` with the argument list of the reference's `utils/lr_finder.py:18`,
the second caller of the train step.  Every trial rate is held for three consecutive batches (`lr_finder.py:27-33`) and each
batch takes one full step through the mirrored module API (forward, calc_loss, backward, clip_grad_norm_, optimizer.step --
the CUDA path).  The reference ends by plotting; here the (rates, losses) it would plot are returned."""
import itertools

import numpy as np

BATCHES_PER_RATE = 3


def lrfind(model, dataloader, optimizer, calc_loss, start=1e-6, stop=4e-3, num_lrs=150, to_screen=False, device="cuda:0"):
    trial_rates = np.repeat(np.logspace(np.log10(start), np.log10(stop), num_lrs), BATCHES_PER_RATE)
    rates, losses = [], []
    for rate, batch in zip(trial_rates, itertools.islice(dataloader, len(trial_rates))):
        x, y, knobs = (t.to(device) for t in batch)
        optimizer.param_groups[0]['lr'] = rate
        y_hat, mag, _ = model.forward(x, knobs)
        # lr_finder.py:38 hands `mag` (not mag_hat) to the loss and no frequency weights; kept as is
        loss = calc_loss(y_hat.float(), y.float(), mag.float())
        rates.append(rate)
        losses.append(loss.item())
        optimizer.zero_grad()
        loss.backward()
        model.clip_grad_norm_()
        optimizer.step()
        if to_screen:
            print(f"lr {rate:.3e}  loss {losses[-1]:.5f}")
    return rates, losses
