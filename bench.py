#!/usr/bin/env python3
"""bench.py -- audio frames/s per train step (comp_4c), BASELINE.json's metric.

  python bench.py --gpus N --steps K --warmup W           # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N ...           # the reference's train step on the host CPU: the unmodified reference
                                                          # when its tree is present (SIGNALTRAIN_REFERENCE, /root/reference), else the
                                                          # numpy port (oracle/st_oracle.py) + the port/reference time ratio measured
                                                          # where both could run (profiles/r02_cpu_port_vs_reference.json)

A "step" is one full iteration of the reference's loop body (train.py:112-151): forward, log-cosh/L1 loss,
backward, L1 clip, Adam, on one batch of synthetic comp_4c windows.  1 audio frame = 1 input PCM sample consumed
(B * chunk per step).  Workload at N=1: BASELINE.json configs[1] (chunk 8192, batch 200,
fp32); at N>1 each rank gets its own batch of 200 windows (weak scaling) and gradients are allreduced over NCCL.

Timing: W warm-up steps, then K steps bracketed by barrier + cuda synchronize, CUDA events on the launching
stream, max over ranks.  Every step reads a different slice of a window pool larger than L2 (config.l2: "pool").
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK_SCALE, SHRINK, KNOBS, SR = 1, 4, 4, 44100
METRIC = "audio frames/sec per train step (comp_4c)"
NCU_SUMMARY = "r02_ae_tm_ncu_full_summary.csv"     # `ncu --set full` of the autoencoder kernels of THIS build (scripts/ncu_summary.py)

# BASELINE.json configs by index.  1 is the bench line (the config the metric is quoted on); 2-4 are the other GPU configs,
# runnable here as extra, clearly labelled lines (--workload N) at N=1 or under torchrun: per-GPU batch as named, synthetic
# windows of the named effect.  "tf32" = the reduced-precision mode (st_set_precision), this repo's answer to the bf16 configs.
WORKLOADS = {
    1: dict(scale=1, knobs=4, batch=200, precision="fp32", effect="Compressor_4c", dtype="f32",
            name="comp_4c synthetic, chunk=8192, batch=200 per GPU, fp32 (BASELINE configs[1])"),
    2: dict(scale=1, knobs=4, batch=512, precision="tf32", effect="Compressor_4c", dtype="tf32",
            name="comp_4c synthetic windows (stand-in for the gen_dataset.py files), chunk=8192, batch=512 per GPU, reduced precision: "
                 "single-pass TF32 products, fp32 storage/accumulate/optimiser (BASELINE configs[2], quoted as bf16)"),
    3: dict(scale=2, knobs=2, batch=256, precision="fp32", effect="Compressor_2knob", dtype="f32",
            name="LA2A-style 2-knob compressor, synthetic, chunk=16384, batch=256 per GPU, fp32 (BASELINE configs[3])"),
    4: dict(scale=1, knobs=1, batch=256, precision="tf32", effect="Denoise", dtype="tf32",
            name="denoise, chunk=8192, batch=256 per GPU, reduced precision: single-pass TF32 products, fp32 "
                 "storage/accumulate/optimiser (BASELINE configs[4], quoted as bf16)"),
}


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line: keep a private handle on the real stdout and point fd 1 at stderr, so that
    anything libraries print there (NCCL's version banner, extension build chatter) cannot land in front of the line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


def measure_tf32_peak(dev):
    """Dense TF32 GEMM rate of this GPU measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3 with TF32
    products allowed, best of 5 after warm-up, CUDA events.  TFLOP/s (2 N^3 flops)."""
    import torch
    n = 8192
    a = torch.randn(n, n, device=dev)
    b = torch.randn(n, n, device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(2):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def ncu_dram_traffic(stage):
    """dram__bytes_read + dram__bytes_write of the stage's kernels from the committed `ncu --set full` summary (bytes per
    stage = sum over its launches), or (None, why)."""
    path = os.path.join(ROOT, "profiles", NCU_SUMMARY)
    want = {"ae_backward": ("ae_bwd_tm_kernel", "ae_track_to_spec_kernel"), "ae_forward": ("ae_fwd_tm_kernel",)}.get(stage)
    if want is None or not os.path.exists(path):
        return None, "no ncu --set full capture of this kernel committed for this build"
    import csv
    rows = list(csv.reader(open(path)))            # scripts/ncu_summary.py: one column per captured launch
    cols = [i for i, n in enumerate(rows[0]) if any(w_ in n for w_ in want)]
    if not cols:
        return None, f"profiles/{NCU_SUMMARY} holds no launch of {want}"
    seen, tot = set(), 0.0
    for i in cols:                                  # one launch of each kernel of the stage
        if rows[0][i] in seen:
            continue
        seen.add(rows[0][i])
        for r in rows[1:]:
            if r[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[i]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(r[1], 1e6)
    return tot, f"profiles/{NCU_SUMMARY} (one launch each of {', '.join(sorted(seen))}, B=200, this round's build)"


def stage_work(d, B):
    """Algorithmic FLOPs / bytes per launch of each stage, SURVEY.md section 8(d) (fp32 words)."""
    BTF, BOF, W = B * d.T * d.F, B * d.OT * d.F, 2 * d.F * d.N
    ae_mac = sum(o * i for o, i in zip([64, 32, 16, 16, 16, 16, 32, 64, d.OT], [d.T, 64, 32, 16, 16 + d.K, 16, 16, 32, 64]))
    p_ae = 2 * (ae_mac + 64 + 32 + 16 * 4 + 32 + 64 + d.OT)
    p_live = 2 * d.F * d.N + 2 * d.N * d.N + p_ae
    return {
        "gemm_analysis": dict(bound="tensor", flops=4 * d.T * d.F * d.N * B, bytes=4 * (B * d.C + W + 2 * BTF)),
        "ae_forward": dict(bound="hbm", flops=4 * d.F * ae_mac * B, bytes=4 * (2 * BTF + BTF + 2 * BOF + 2 * BOF + B * d.K + p_ae)),
        "gemm_synthesis": dict(bound="tensor", flops=4 * d.OT * d.F * d.N * B, bytes=4 * (2 * BOF + W + B * d.OT * d.N)),
        "gemm_synthesis_dgrad": dict(bound="tensor", flops=4 * d.OT * d.F * d.N * B, bytes=4 * (B * d.L + 2 * BOF + W)),
        "gemm_synthesis_wgrad": dict(bound="tensor", flops=4 * d.OT * d.F * d.N * B, bytes=4 * (B * d.L + 2 * BOF + W)),
        "ae_backward": dict(bound="hbm", flops=12 * d.F * ae_mac * B, bytes=4 * (2 * BOF + 4 * BTF + 2 * p_ae)),
        "gemm_analysis_wgrad": dict(bound="tensor", flops=4 * d.T * d.F * d.N * B, bytes=4 * (B * d.C + 2 * BTF + W)),
        "adam": dict(bound="hbm", flops=12 * p_live, bytes=7 * 4 * p_live),
        "pack_weights": dict(bound="hbm", flops=0, bytes=4 * (4 * d.N * d.N * 3 // 4 + 2 * W)),
        "finalize_dft_grads": dict(bound="hbm", flops=0, bytes=4 * (2 * W + 4 * d.N * d.N)),
        "l1_norm": dict(bound="hbm", flops=0, bytes=4 * 4 * d.N * d.N),
        "overlap_add": dict(bound="hbm", flops=0, bytes=4 * (B * d.OT * d.N + 2 * B * d.L)),
        "loss": dict(bound="hbm", flops=0, bytes=4 * (3 * B * d.L + 2 * BOF)),
    }


class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def workload_name(wl, B):
    """The SAME string in both arms (the GPU count lives in n_gpus / config.gpus)."""
    return wl["name"].replace(f"batch={wl['batch']} per GPU", f"batch={B} per GPU")


def port_over_reference():
    """Time ratio port / unmodified reference measured where both could run (scripts/measure_port_vs_reference.py)."""
    path = os.path.join(ROOT, "profiles", "r02_cpu_port_vs_reference.json")
    if os.path.exists(path):
        j = json.load(open(path))
        return float(j["port_over_reference_time"]), f"profiles/r02_cpu_port_vs_reference.json (B={j['batch']}, {j['threads']} threads, {j['where']})"
    return None, "not measured"


def cpu_reference_step_rate(B, steps, warmup, threads, wl=None, batches=None, params=None):
    """The reference's algorithm (oracle port, float32) on the host cores: frames/s, ms/step and the per-step losses.
    batches: optional list of (x, y, knobs) numpy batches to step through in order (the oracle replay of bench.py's first
    steps); params: optional initial parameters (name -> array), default the port's own seeded init."""
    wl = wl or WORKLOADS[1]
    from oracle import st_oracle as O
    from signaltrain_b200 import data
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    d = O.model_dims(wl["scale"], SHRINK, wl["knobs"])
    if batches is None:
        batches = [data.make_pool(B, d.C, d.L, getattr(data, wl["effect"])(), SR, seed=218)]
    lr_sched, _ = O.get_1cycle_schedule(1e-4, 200000, 1000, 200)
    tr = O.Trainer(d, params if params is not None else O.init_params(d, seed=218), lr_sched, dtype=np.float32)
    ctx = threadpool_limits(limits=threads) if threadpool_limits else None
    losses = []
    t0 = time.perf_counter()
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        x, y, k = batches[i % len(batches)]
        losses.append(tr.step(x, y, k)[0])
    dt = time.perf_counter() - t0
    if ctx is not None:
        ctx.unregister() if hasattr(ctx, "unregister") else None
    return B * d.C * steps / dt, 1e3 * dt / steps, losses


def run_reference(args):
    """The reference arm: the reference's own CPU train step on this box's host cores, all threads, same workload string, metric
    and unit as the native arm.  Each step is one full train step of the workload's batch (about 1-2 s of CPU work), so the
    requested --steps/--warmup are honoured up to a budget of ~150 s of CPU time; a clamp is stated in the line."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = WORKLOADS[args.workload]
    B = args.batch or wl["batch"]
    from oracle import ref_loader
    from oracle import st_oracle as O
    from signaltrain_b200 import data
    ref = ref_loader.load_reference()
    d = O.model_dims(wl["scale"], SHRINK, wl["knobs"])
    x, y, k = data.make_pool(B, d.C, d.L, getattr(data, wl["effect"])(), SR, seed=218)
    run = (lambda st_, w_: ref_loader.reference_step_rate(ref, x, y, k, st_, w_, threads, scale=wl["scale"], shrink=SHRINK)) if ref is not None \
        else (lambda st_, w_: cpu_reference_step_rate(B, st_, w_, threads, wl)[:2])
    _, probe_ms = run(1, 1)                                       # one warm step to size the run
    budget_steps = max(2, int(150e3 / max(probe_ms, 1.0)))
    warm = max(0, min(args.warmup, budget_steps // 4))
    steps = max(1, min(args.steps, budget_steps - warm))
    fps, ms = run(steps, warm)
    ratio, ratio_src = port_over_reference()
    kind = "reference" if ref is not None else "port"
    cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
           "sample": f"{steps} full train steps of B={B} windows (+{warm + 2} warm-up) by " +
                     ("the unmodified reference (signaltrain.nn_proc.st_model + calc_loss + clip_grad_norm_ + torch.optim.Adam, torch CPU, "
                      f"{threads} threads)" if ref is not None else f"oracle/st_oracle.py Trainer (float32 numpy/BLAS, {threads} threads)")}
    if ref is None:
        cpu["port_over_reference_time"] = ratio
        cpu["port_over_reference_source"] = ratio_src
        cpu["note"] = ("the reference tree does not travel to this box: this is the numpy port; where both ran, the port took "
                       f"{ratio:.2f}x the unmodified reference's time per step" if ratio else "the reference tree does not travel to this box: numpy port")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(wl, B), "where": "host CPU (float32)", "gpus": args.gpus,
                       "steps_requested": args.steps, "warmup_requested": args.warmup,
                       "steps_clamped": steps != args.steps or warm != args.warmup},
            "cpu_baseline": cpu,
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_native(args):
    import torch
    import torch.distributed as dist
    import signaltrain_b200 as st
    from signaltrain_b200 import data
    from signaltrain_b200.train import FusedTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the measured path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from signaltrain_b200 import parallel as _parallel
        _parallel.nccl_env_defaults(world)
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    CHUNK_SCALE, KNOBS = wl["scale"], wl["knobs"]
    B, K, Wm = args.batch or wl["batch"], args.steps, max(3, args.warmup)

    torch.manual_seed(218)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # stdout carries exactly one JSON line
        model = st.nn_proc.st_model(scale_factor=CHUNK_SCALE, shrink_factor=SHRINK, num_knobs=KNOBS, sr=SR).to(dev)
    model.set_precision(wl["precision"])
    C, L = model.in_chunk_size, model.out_chunk_size
    lr_sched, _ = st.learningrate.get_1cycle_schedule(lr_max=1e-4, n_data_points=200000, epochs=1000, batch_size=200)
    trainer = FusedTrainer(model, lr_sched)
    # every measured launch goes to ONE explicit stream (the legacy default stream cannot be captured into a CUDA graph, which
    # st_train_step does with its ~17 launches after it has seen the same buffers twice); events are recorded on that stream
    bench_stream = torch.cuda.Stream(device=dev)
    bench_stream.wait_stream(torch.cuda.current_stream(dev))
    torch.cuda.set_stream(bench_stream)

    # window pool larger than L2 (126 MB): P windows * (C + L + K) * 4 B
    nbatch = max(4, -(-int(168e6) // (B * (C + L + KNOBS) * 4)))
    P = nbatch * B
    xh, yh, kh = data.make_pool(P, C, L, getattr(data, wl["effect"])(), SR, seed=218 + rank)
    xp, yp, kp = (torch.from_numpy(a).pin_memory() for a in (xh, yh, kh))
    xd, yd, kd = xp.to(dev), yp.to(dev), kp.to(dev)

    def batch_of(i, src):
        s = (i % nbatch) * B
        return src[0][s:s + B], src[1][s:s + B], src[2][s:s + B]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the first steps are replayed by the oracle in the cpu_baseline leg: keep the initial parameters and the losses
    replay_n = 3 if (world == 1 and not args.no_cpu_baseline) else 0
    P0 = {n: p.detach().cpu().numpy().copy() for n, p in zip([n for n, _ in model.named_parameters()], model.parameters())} if replay_n else None
    first_losses = []
    it = 0
    for _ in range(Wm):
        l_ = trainer.step(*batch_of(it, (xd, yd, kd)))
        if it < replay_n:
            first_losses.append(l_.clone())
        it += 1
    eng = trainer.eng
    # data-parallel replicas must hold identical parameters (deterministic update of identically reduced gradients)
    replicas_equal = None
    if world > 1:
        chk = torch.stack([p.detach().double().sum() for p in trainer.params] + [p.detach().double().abs().sum() for p in trainer.params])
        allchk = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        replicas_equal = all(bool(torch.equal(allchk[0], c)) for c in allchk[1:])
    tf32_peak = measure_tf32_peak(dev) if rank == 0 else None
    # ---- timed region: K steps, device-resident inputs ------------------------------------------------------
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        loss = trainer.step(*batch_of(it, (xd, yd, kd)))
        it += 1
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0 + (K if world > 1 else 0)
    clk = clocks.stop() if rank == 0 else None
    final_loss = float(loss.item())
    # ---- end-to-end: host (pinned) inputs, H2D copies and a D2H loss read inside the timed region ------------
    Ke = K
    host_batches = [batch_of(it + i, (xp, yp, kp)) for i in range(Ke)]          # views of the pinned pool
    it += Ke
    trainer.run_host_batches([batch_of(it + i, (xp, yp, kp)) for i in range(3)])   # warm-up: stream, buffers, pinned slots
    it += 3
    barrier()
    t0 = time.perf_counter()
    e2e_losses = trainer.run_host_batches(host_batches)       # per step: H2D of x, y, knobs; the step; D2H of the loss
    barrier()
    e2e_s = time.perf_counter() - t0
    assert len(e2e_losses) == Ke
    h2d = B * (C + L + KNOBS) * 4
    # ---- per-stage device time (separate pass: the event pairs perturb the pipeline slightly) ---------------
    stages = {}
    eng.profile(rank == 0)
    if rank == 0:
        eng.profile_read()
    for _ in range(max(3, min(K, 10))):          # every rank steps (the allreduce is collective); rank 0 records events
        trainer.step(*batch_of(it, (xd, yd, kd)))
        it += 1
    if rank == 0:
        stages = {k: (ms / max(c, 1), c) for k, (ms, c) in eng.profile_read().items() if c}
    eng.profile(False)
    # ---- max over ranks ---------------------------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    d = eng.g                          # geometry of the measured handle (C, N, H, T, OT, L, F, K)
    peaks = measured_peaks()
    work = stage_work(d, B)
    step_sum = sum(v[0] for v in stages.values())
    top = max((k for k in stages if k in work), key=lambda k: stages[k][0])
    w = work[top]
    dur_s = stages[top][0] * 1e-3
    if w["bound"] == "tensor":       # fp32-fidelity GEMMs are 3xTF32: algorithmic FLOPs against the measured TF32 rate
        achieved, peak, unit = w["flops"] / dur_s / 1e12, tf32_peak, "TFLOP/s"
    else:
        achieved, peak, unit = w["bytes"] / dur_s / 1e9, peaks["hbm"], "GB/s"
    traffic, traffic_src = ncu_dram_traffic(top) if args.workload == 1 and B == 200 else (None, "no ncu --set full capture at this workload")
    roofline = {"kernel": top, "bound": w["bound"], "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peaks["src"], "kernel_ms": stages[top][0],
                "share_of_step": stages[top][0] / step_sum if step_sum else None,
                "algorithmic_bytes": w["bytes"], "algorithmic_flops": w["flops"],
                "stages_ms": {k: round(v[0], 4) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])}}
    roofline["tf32_peak_tflops"] = tf32_peak
    roofline["tf32_peak_source"] = "measured in this run: torch.matmul 8192^3 with TF32 products, best of 5 (same recipe as MEASURED_PEAKS.json's bf16)"
    gemm_flops = sum(work[k]["flops"] for k in stages if k.startswith("gemm_") and k in work)
    gemm_ms = sum(stages[k][0] for k in stages if k.startswith("gemm_") and k in work)
    if gemm_ms > 0:
        roofline["gemms"] = {"ms": gemm_ms, "algorithmic_tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12,
                             "frac_of_tf32_peak": gemm_flops / (gemm_ms * 1e-3) / 1e12 / tf32_peak,
                             "note": "3xTF32: the tensor pipe issues 3x the algorithmic FLOPs"}
    # CPU baseline: the oracle port on this box's host cores, bounded sample
    cores = os.cpu_count() or 1
    cpu_fps, cpu_ms, replay = (None, None, None)
    if world == 1 and not args.no_cpu_baseline:
        # the port steps through THE SAME first batches from THE SAME initial parameters: timing sample + oracle replay
        host_first = [tuple(a_[(i % nbatch) * B:(i % nbatch) * B + B] for a_ in (xh, yh, kh)) for i in range(replay_n)]
        cpu_fps, cpu_ms, ol = cpu_reference_step_rate(B, replay_n - 1, 1, cores, wl, batches=host_first, params=P0)
        gl = [float(l_.item()) for l_ in first_losses]
        replay = {"steps": replay_n, "gpu_losses": gl, "oracle_losses": ol,
                  "max_abs_diff": max(abs(a_ - b_) for a_, b_ in zip(gl, ol)),
                  "note": "first train steps of this run replayed by oracle.Trainer (float32) from the same initial parameters on the same batches"}
    frames = world * B * C
    line = {"metric": METRIC, "value": frames * K / (ms_total * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": workload_name(wl, B), "gpus": world, "precision": eng.precision,
                       "global_batch": world * B, "windows_per_s": world * B * K / (ms_total * 1e-3),
                       "stft_frames_per_s": world * B * d.T * K / (ms_total * 1e-3),
                       "l2": f"pool: each step reads a different batch of a {P}-window ({P * (C + L + KNOBS) * 4 / 1e6:.0f} MB) pool",
                       "parallelism": f"dp{world}", "final_loss": final_loss},
            "e2e": {"value": frames * Ke / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * e2e_s / Ke,
                    "api": "signaltrain_b200.train.FusedTrainer.run_host_batches: pinned host batches, H2D of batch i+1 on a copy stream while step i (st_train_step) runs, every step's loss read back to the host (one step late)"},
            "gpu_launches": int(launches), "cuda_graph_replays": eng.graph_replays(), "simt_fallbacks": eng.fallback_count(),
            "clocks": clk, "roofline": roofline}
    if world > 1:
        line["config"]["exchange"] = (f"{os.environ.get('ST_DP_EXCHANGE', 'packed')}: one NCCL allreduce of the packed gradient payload per step "
                                      f"(NCCL_ALGO={os.environ.get('NCCL_ALGO', 'default')}, NCCL_PROTO={os.environ.get('NCCL_PROTO', 'default')})")
    if replicas_equal is not None:
        line["replica_parameters_identical"] = replicas_equal
    if cpu_fps is not None:
        ratio, ratio_src = port_over_reference()
        line["cpu_baseline"] = {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port", "ms_per_step": cpu_ms,
                                "sample": f"{replay_n - 1} full train steps of B={B} windows (+1 warm-up) by oracle/st_oracle.py (float32 numpy/BLAS, "
                                          f"{cores} threads), on this run's first batches",
                                "port_over_reference_time": ratio, "port_over_reference_source": ratio_src}
        line["oracle_replay"] = replay
    if line["simt_fallbacks"] != 0:
        raise RuntimeError(f"bench.py: {line['simt_fallbacks']} calls of the timed path fell back to SIMT kernels")
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="windows per GPU per step (default: the workload's)")
    ap.add_argument("--workload", type=int, default=1, choices=sorted(WORKLOADS),
                    help="index into BASELINE.json configs (1 = the bench line; 2-4 = the other GPU configs, extra lines)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
