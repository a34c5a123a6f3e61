"""Host-side mirror of `signaltrain/train.py`: `train()` (:167-278), `train_loop()` (:84-164) and
`eval_status_save()` (:28-80), re-hosted on the fused CUDA train step.

`FusedTrainer.step` is one iteration of the loop body train.py:104-151 -- forward, calc_loss, backward, L1 clip,
Adam, and the one-step learning-rate lag -- as a single C-ABI call (or forward/backward + NCCL allreduce + Adam
when data-parallel).  The orchestration around it (epochs, status line, validation EMA, log files, checkpoints)
follows the reference's behaviour; plots are out of scope."""
import os
import time

import numpy as np
import torch

from . import data as st_data
from . import learningrate, loss_functions, misc, nn_proc, optim, parallel


class FusedTrainer:
    """model + Adam state + lr schedule, stepping through engine.train_step.

    Data parallel (world_size > 1): every rank holds a replica and its own windows; gradients are summed with one
    NCCL allreduce over a flat buffer, scaled by 1/world inside the clip+Adam launch, so replicas stay identical
    (SURVEY.md section 8e)."""

    def __init__(self, model, lr_sched, optimizer=None, l1_lambda=2e-5, process_group=None, distributed=None):
        self.model = model
        self.params = [p.data for p in model.ordered_parameters()]
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedTrainer: move the model to a CUDA device first (no CPU fallback)")
        self.device = dev
        self.optimizer = optimizer if optimizer is not None else optim.Adam(model, lr=float(lr_sched[0]))
        _, self.m, self.v = self.optimizer._state_lists()
        # one flat gradient buffer (16-byte aligned slots) so a single collective covers all 40 tensors
        fb = parallel.FlatBuffer([tuple(p.shape) for p in self.params], dev)
        self.flat_grads, self.grads = fb.flat, fb.views
        self._fb = fb
        self.reducer = None
        for p, g in zip(model.ordered_parameters(), self.grads):
            p.grad = g
        self.lr_sched = np.asarray(lr_sched, dtype=np.float64)
        self.lr = float(self.lr_sched[0])
        self.iter_count = 0
        self.l1_lambda = l1_lambda
        self.sbf = None
        self.eng = None
        self.pg = process_group
        self.world = parallel.world_size(self.pg) if distributed is not False else 1
        if self.world > 1:
            parallel.broadcast_parameters(self.params, 0, self.pg)
        # two loss slots, used alternately: the host loop reads step i's loss from a side stream while step i + 1 writes the other
        self._loss_bufs = [torch.zeros((), device=dev, dtype=torch.float32) for _ in range(2)]
        self.loss_buf = self._loss_bufs[0]

    def _setup(self, x):
        self.eng = self.model.mpaec._engine_for(x)
        F = self.eng.g.F
        # train.py:115-117: exp(7/F * arange(F)) in float32
        self.sbf = torch.exp(torch.tensor(7.0 / F, dtype=torch.float32) * torch.arange(0., F)).float().to(self.device)

    def step(self, x, y, knobs):
        """Returns the (device, 0-dim) loss of this batch.  Asynchronous: nothing here syncs the host."""
        if self.eng is None:
            self._setup(x)
        eng = self.eng
        step_no = self.optimizer._step + 1
        self.loss_buf = self._loss_bufs[step_no & 1]
        if self.world == 1:
            hp = eng.adam_hp(self.lr, step_no, max_norm=1.0)
            eng.train_step(x, y, knobs, self.params, self.grads, self.m, self.v, self.sbf, self.l1_lambda / 10, hp,
                           loss_out=self.loss_buf)
        else:
            mode = os.environ.get("ST_DP_EXCHANGE", "packed")
            if mode in ("overlap", "sliced"):
                y_hat, _, mag_hat, _ = eng.forward(x, knobs, self.params)
                loss, g_y, g_m = eng.loss(y_hat, y, mag_hat, self.sbf, self.l1_lambda / 10)
                if self.reducer is None:
                    self.reducer = parallel.GradReducer(self._fb, [tuple(p.shape) for p in self.params], eng.g.F, self.pg)
            if mode == "overlap":                 # synthesis pair reduced while the rest of the backward runs
                eng.backward(g_y, None, g_m, self.params, self.grads, part="begin")
                self.reducer.start_synthesis()
                eng.backward(g_y, None, g_m, self.params, self.grads, part="finish")
                scale = self.reducer.finish()
            elif mode == "sliced":                # after the backward, only the rows that carry gradient (12.6 MB, 4 calls)
                eng.backward(g_y, None, g_m, self.params, self.grads)
                self.reducer.start_synthesis()
                scale = self.reducer.finish()
            elif mode == "packed":                # default: st_grad_step, then ONE allreduce of the packed payload (8.5 MB: analysis
                # rows >= F are zero and the synthesis pair is Hermitian), scattered back by st_unpack_grads
                # (the gradients leave the backward AS that payload: st_grad_step_packed), and st_unpack_clip computes the L1 clip
                # coefficient while it scatters, so the update below is the live-row Adam launch of the single-GPU step
                if getattr(self, "_packed", None) is None:
                    self._packed = torch.empty(eng.packed_grad_floats(), device=self.device, dtype=torch.float32)
                loss = eng.grad_step_packed(x, y, knobs, self.params, self._packed, self.sbf, self.l1_lambda / 10, loss_out=self.loss_buf)
                scale = parallel.allreduce_sum_(self._packed, self.pg)
                eng.unpack_clip(self._packed, self.grads, scale, 1.0)
                hp = eng.adam_hp(self.lr, step_no, grad_scale=scale, max_norm=1.0)
                eng.adam_step_clipped(self.params, self.grads, self.m, self.v, hp)
                self.loss_buf = loss
                self.optimizer._step = step_no
                self.lr = float(self.lr_sched[min(self.iter_count, len(self.lr_sched) - 1)])
                self.optimizer.param_groups[0]['lr'] = self.lr
                self.iter_count += 1
                return self.loss_buf
            else:                                 # "whole": st_grad_step (forward + fused loss tail + backward in one call),
                # then ONE allreduce of the whole flat buffer (16.8 MB).  Measured on 2 B200s (ms/step, v7): whole 1.011,
                # sliced 1.056, overlap 1.060 -- at this size NCCL is launch / latency bound, so four smaller collectives
                # cost more than the 25 % of payload they save, and the overlapped one cannot get SMs while the
                # autoencoder backward holds them all
                loss = eng.grad_step(x, y, knobs, self.params, self.grads, self.sbf, self.l1_lambda / 10, loss_out=self.loss_buf)
                scale = parallel.allreduce_sum_(self.flat_grads, self.pg)
            hp = eng.adam_hp(self.lr, step_no, grad_scale=scale, max_norm=1.0)
            eng.adam_step(self.params, self.grads, self.m, self.v, hp)
            self.loss_buf = loss
        self.optimizer._step = step_no
        # train.py:150: the schedule value of THIS iteration is installed after the update (one-step lag)
        self.lr = float(self.lr_sched[min(self.iter_count, len(self.lr_sched) - 1)])
        self.optimizer.param_groups[0]['lr'] = self.lr
        self.iter_count += 1
        return self.loss_buf

    def sync_optimizer_state(self):
        for p in self.model.ordered_parameters():
            self.optimizer.state[p]["step"] = torch.tensor(float(self.optimizer._step))

    def run_host_batches(self, batches, on_loss=None):
        """Train on an iterable of pinned HOST batches (x, y, knobs) -- what `for x, y, knobs in dataloader` feeds the
        reference's train_loop (train.py:104-106) with pin_memory=True.  The host->device copy of batch i+1 runs on a
        copy stream while step i computes (two device buffers), and every step's loss is read back to the host one step
        late (the reference reads it every 10th batch, train.py:125-129), so the host never waits for the step it has just
        launched.  Returns the list of per-step losses (python floats); on_loss(i, value) is called as they arrive."""
        dev = self.device
        if getattr(self, "_feed", None) is None:             # created once: stream, events, pinned loss slots, device buffers
            self._feed = {"stream": torch.cuda.Stream(device=dev), "bufs": [None, None],
                          "ready": [torch.cuda.Event(), torch.cuda.Event()], "free": [torch.cuda.Event(), torch.cuda.Event()],
                          "loss_host": [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)],
                          "loss_ev": [torch.cuda.Event(), torch.cuda.Event()]}
        fd = self._feed
        copy_stream, bufs, ready, free = fd["stream"], fd["bufs"], fd["ready"], fd["free"]
        loss_host, loss_ev = fd["loss_host"], fd["loss_ev"]
        main = torch.cuda.current_stream(dev)
        losses = []

        def upload(slot, batch):
            with torch.cuda.stream(copy_stream):
                if bufs[slot] is None or any(b.shape != t.shape for b, t in zip(bufs[slot], batch)):
                    bufs[slot] = tuple(torch.empty(t.shape, dtype=torch.float32, device=dev) for t in batch)
                    copy_stream.wait_stream(main)
                else:
                    copy_stream.wait_event(free[slot])          # the step that read this buffer has finished
                for dst, src in zip(bufs[slot], batch):
                    dst.copy_(src, non_blocking=True)
                ready[slot].record(copy_stream)

        def collect(i):
            loss_ev[i % 2].synchronize()
            v = float(loss_host[i % 2])
            losses.append(v)
            if on_loss is not None:
                on_loss(i, v)

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return losses
        upload(0, nxt)
        i = 0
        while nxt is not None:
            cur = i % 2
            nxt = next(it, None)
            if nxt is not None:
                upload(1 - cur, nxt)
            main.wait_event(ready[cur])
            if i >= 2:
                main.wait_event(loss_ev[cur])                # the read-back of the loss slot this step writes (step i - 2's) is done
            loss = self.step(*bufs[cur])
            free[cur].record(main)
            # loss read-back on the copy stream: a 4-byte D2H in the step's own stream would put the copy engine's latency
            # between every two steps
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[cur])
                loss_host[cur].copy_(loss, non_blocking=True)
                loss_ev[cur].record(copy_stream)
            if i > 0:
                collect(i - 1)
            i += 1
        collect(i - 1)
        return losses


def eval_status_save(model, effect, epoch, epochs, lr, mom, device, dataloader_val, logfilename, first_time,
                     beta, vl_avg, out_checkpointname, parallel, optimizer, data_point, smoothed_loss, y_size, sr,
                     status_every, plot_every=10, cp_every=25, scale_by_freq=None):
    """Validation pass, log files, checkpoint (train.py:28-80).  Plot calls are out of scope."""
    model.eval()
    val_batch_num = 0
    y_val_hat = y_val_cuda = None
    with torch.no_grad():
        for x_val, y_val, knobs_val in dataloader_val:
            val_batch_num += 1
            x_val_cuda, y_val_cuda, knobs_val_cuda = x_val.to(device), y_val.to(device).float(), knobs_val.to(device)
            y_val_hat, mag_val, mag_val_hat = model.forward(x_val_cuda, knobs_val_cuda)
            loss_val = loss_functions.calc_loss(y_val_hat, y_val_cuda, mag_val_hat, scale_by_freq=scale_by_freq)
            vl_avg = beta * vl_avg + (1 - beta) * loss_val.item()
            if 0 == val_batch_num % status_every:
                print(f"\repoch {epoch+1}/{epochs}, time: {time.time() - first_time:.2f}: lr={lr:.2e},mom={mom:.3f} "
                      f"data_point {data_point}: loss: {smoothed_loss:.3e} val_loss: {vl_avg:.3e}   ", end="")
    with open(logfilename, "a") as f:
        f.write(f"{epoch+1} {vl_avg:.3e}\n")
    if y_val_hat is not None:
        with open("val_err_mae.dat", "a") as f:
            f.write(f"{epoch+1} {loss_functions.mae(y_val_hat, y_val_cuda).item():.3e}\n")
    if ((epoch + 1) % cp_every == 0) or (epoch == epochs - 1):
        misc.save_checkpoint(out_checkpointname, model, epoch, parallel, optimizer, effect, sr)
    if (epoch + 1) == 1:
        hours = (time.time() - first_time) * (epochs - 1) / 3600.0
        print(f"\nExpect run to finish in roughly {hours:.1f} hours")
    return vl_avg


def train_loop(model, effect, device, optimizer, epochs, batch_size, lr_sched, mom_sched, dataloader, dataloader_val,
               y_size, parallel, logfilename, out_checkpointname, plot_every=10, cp_every=25, sr=44100, lr_max=1e-4):
    """The training loop of train.py:84-164 on the fused step: same status line (EMA of every 10th batch's loss,
    beta 0.98, bias-corrected), same lr/momentum bookkeeping, validation + checkpoint per epoch."""
    trainer = FusedTrainer(model, lr_sched, optimizer=optimizer)
    batch_num, status_every = 0, 10
    avg_loss, vl_avg, beta = 0.0, 0.0, 0.98
    smoothed_loss = float("nan")
    first_time = time.time()
    lr, mom = float(lr_sched[0]), float(mom_sched[0])
    for epoch in range(epochs):
        print("")
        data_point = 0
        model.train()
        for x, y, knobs in dataloader:
            x_cuda = x.to(device, non_blocking=True).float()
            y_cuda = y.to(device, non_blocking=True).float()          # y may arrive as float64 (train.py:120)
            knobs_cuda = knobs.to(device, non_blocking=True).float()
            lr = lr_sched[min(trainer.iter_count, len(lr_sched) - 1)]
            mom = mom_sched[min(trainer.iter_count, len(mom_sched) - 1)]
            data_point += batch_size
            loss = trainer.step(x_cuda, y_cuda, knobs_cuda)
            optimizer.param_groups[0]['momentum'] = mom               # ignored by Adam, kept for parity (train.py:151)
            batch_num += 1
            if 0 == batch_num % status_every:                         # the only host sync in the loop (train.py:125-129)
                avg_loss = beta * avg_loss + (1 - beta) * loss.item()
                smoothed_loss = avg_loss / (1 - beta ** batch_num)
                print(f"\repoch {epoch+1}/{epochs}, time: {time.time() - first_time:.2f}: lr={lr:.2e},mom={mom:.3f}, "
                      f"data_point {data_point}: loss: {smoothed_loss:.3e}   ", end="")
        trainer.sync_optimizer_state()
        sbf = trainer.sbf
        vl_avg = eval_status_save(model, effect, epoch, epochs, lr, mom, device, dataloader_val, logfilename, first_time,
                                  beta, vl_avg, out_checkpointname, parallel, optimizer, data_point, smoothed_loss, y_size,
                                  sr, status_every, scale_by_freq=sbf)
    print("\nTotal elapsed time for training loop =", time.time() - first_time)
    return None


def train(effect=None, epochs=100, n_data_points=200000, batch_size=20, device=None, plot_every=10, cp_every=25,
          sr=44100, datapath=None, scale_factor=1, shrink_factor=4, apex_opt="O0", target_type="stream", lr_max=1e-4,
          in_checkpointname='modelcheckpoint.tar', compand=False):
    """Main training routine; arguments as in the reference (train.py:167-170).  Returns the trained model."""
    if effect is None:
        effect = st_data.Compressor_4c()
    if device is None:
        device = torch.device("cuda:0")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("signaltrain_b200.train: this train step runs on a CUDA (B200) device only; "
                           "there is no CPU fallback")
    # the reference's only precision knob (train.py:133-136,169,184): "O0" = fp32; "O1".."O3" = apex mixed precision.
    # Here: "O0" -> fp32-faithful products (3xTF32); anything else -> single-pass TF32 products, fp32 everything else.
    precision = "fp32" if apex_opt in (None, "O0") else "tf32"
    if precision != "fp32":
        print(f"*** NOTE: apex_opt={apex_opt}: reduced-precision mode (single-pass TF32 tensor-core products; fp32 storage, "
              "accumulation, loss and optimiser)")
    print(f'SignalTrain (B200) training began at {time.ctime()}. Options:')
    print(f'    epochs = {epochs}, n_data_points = {n_data_points}, batch_size = {batch_size}')
    print(f'    scale_factor = {scale_factor}, shrink_factor = {shrink_factor}')
    num_knobs = len(effect.knob_names)
    print(f'    num_knobs = {num_knobs}')
    effect.info()

    state_dict, rv = misc.load_checkpoint(in_checkpointname, fatal=False, device="cpu")
    if state_dict != {}:
        scale_factor, shrink_factor = rv['scale_factor'], rv['shrink_factor']
        sr = rv['sr']
    model = nn_proc.st_model(scale_factor=scale_factor, shrink_factor=shrink_factor, num_knobs=num_knobs, sr=sr)
    model.set_precision(precision)
    if state_dict != {}:
        model.load_state_dict(state_dict)
    chunk_size, out_chunk_size = model.in_chunk_size, model.out_chunk_size
    y_size = out_chunk_size
    print("Model defined.  Number of trainable parameters:", sum(p.numel() for p in model.parameters() if p.requires_grad))
    print("      model.in_chunk_size, model.out_chunk_size = ", model.in_chunk_size, model.out_chunk_size)
    model.to(device)

    lr_sched, mom_sched = learningrate.get_1cycle_schedule(lr_max=lr_max, n_data_points=n_data_points, epochs=epochs,
                                                           batch_size=batch_size)
    optimizer = optim.Adam(model, lr=lr_sched[0], weight_decay=0)
    if state_dict != {} and 'optimizer' in rv and rv['optimizer'].get('state'):
        try:      # the reference leaves this as a TODO (train.py:229); resuming Adam state is strictly better
            optimizer.load_state_dict(rv['optimizer'])
        except Exception as e:   # pragma: no cover
            print("    (optimizer state in checkpoint not restored:", e, ")")

    if datapath is None:
        if not hasattr(effect, "apply"):
            raise RuntimeError("on-the-fly synthesis needs a signaltrain_b200.data effect; for the reference's own "
                               "Effect classes build the DataLoaders with the reference's datasets.py and call train_loop()")
        dataloader = st_data.SynthWindowBatches(chunk_size, effect, sr=sr, datapoints=n_data_points, batch_size=batch_size,
                                                y_size=out_chunk_size, augment=True)
        dataloader_val = st_data.SynthWindowBatches(chunk_size, effect, sr=sr, datapoints=max(batch_size, n_data_points // 4),
                                                    batch_size=batch_size, y_size=out_chunk_size, augment=False,
                                                    recycle=True, seed=99991)
    else:
        # prerecorded input / target files (train.py:240-246): the corpus is preloaded into HBM, windows are cropped (and, for
        # target_type != "stream", the effect re-run) on the device; the host only draws the reference's random numbers
        from . import device_data
        rerun = target_type != "stream"
        dataloader = device_data.file_batches(datapath + "/Train/", effect, chunk_size, out_chunk_size, batch_size, n_data_points,
                                              device, sr=sr, rerun=rerun, augment=True, compand=compand)
        dataloader_val = device_data.file_batches(datapath + "/Val/", effect, chunk_size, out_chunk_size, batch_size,
                                                  max(batch_size, n_data_points // 4), device, sr=sr, rerun=rerun, augment=False,
                                                  compand=compand)

    logfilename = "vl_avg_out.dat"
    open(logfilename, "a").close()
    out_checkpointname = "modelcheckpoint.tar"
    train_loop(model, effect, device, optimizer, epochs, batch_size, lr_sched, mom_sched, dataloader, dataloader_val,
               y_size, False, logfilename, out_checkpointname, sr=sr, lr_max=lr_max)
    return model
