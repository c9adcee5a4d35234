#!/usr/bin/env python
"""bench.py -- headline benchmark of the RAVE waveform hot path on B200.

    python bench.py --gpus N --steps K --warmup W          # our arm (N>1 under torchrun)
    python bench.py --impl reference ...                   # the reference's CPU arithmetic (oracle port)

Metric (BASELINE.json): audio-seconds/s of the v2 @48 kHz training step (PQMF + encoder + generator
+ v2 discriminator, forward + backward + optimiser), B = 32 x 65536 per GPU, synthetic data.
One "step" = one `RAVE.training_step` in phase 2, following the reference's schedule (1 D-step every
`update_discriminator_every` = 4 batches, rave/configs/v2.gin:86): the timed K steps always cover
whole 4-step cycles' worth of alternation starting at batch_idx 0.

Prints ONE JSON line (rank 0).  Keys: see the task contract; `roofline` is for the dominant kernel
(timed alone, CUDA events, inputs > L2 or L2 flushed), `cpu_baseline` is the oracle port on the
host cores for a bounded sample, `e2e` goes through the public API with pinned host buffers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
T = 65536


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (BASELINE: 32)")
    ap.add_argument("--config", default="v2")
    ap.add_argument("--precision", default=os.environ.get("RAVE_B200_PRECISION", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="issue every launch from Python (no CUDA graphs)")
    ap.add_argument("--cudnn-baseline", action="store_true",
                    help="also time the SAME step arithmetic through stock torch/cuDNN (TF32, as scripts/train.py:135-136 "
                         "configures the reference) on this GPU and report it as `stock_cudnn_tf32`")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md): NVML, else nvidia-smi."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def _run_nvml(self):
        """NVML samples every ~5 ms (a timed region of 8 steps is ~100 ms: nvidia-smi is too slow for that)."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self._stop_evt.is_set():
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for b, n in bits.items():
                if r & b:
                    self.reasons.add(n)
            self._stop_evt.wait(0.005)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(s))


def synthetic_batch(B, seed=1234):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (0.5 * torch.randn(B, 1, T, generator=g)).clamp(-1, 1)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's arithmetic on the host cores
# ------------------------------------------------------------------------------------------------

def cpu_reference_run(args, steps, warmup, budget_s):
    """Times fwd+bwd of the phase-2 step (oracle port of rave/model.py:288-424) on the CPU for a
    bounded sample: B_cpu x 65536 per step, alternating D/G like the GPU arm."""
    import torch
    from oracle import rave_oracle as O
    from rave_b200 import configs
    # oneDNN convs of this size stop scaling (and regress) far below 128 threads: use one socket's worth
    cores = min(os.cpu_count() or 1, int(os.environ.get("RAVE_CPU_THREADS", "32")))
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    m = configs.build_rave(args.config, sampling_rate=SR)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    del m
    cfg = O.ArchConfig() if args.config == "v2" else O.v2_small_config()
    B_cpu = 2
    x = synthetic_batch(B_cpu)
    eps = torch.randn(B_cpu, cfg.latent_size, T // (16 * int(__import__("numpy").prod(cfg.ratios))))
    times = []
    t_start = time.time()
    n = 0
    for i in range(warmup + steps):
        t0 = time.time()
        O.train_step_cpu(x, sd, cfg, eps, dis_step=(i % 4 == 0))
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        n += 1
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    mean = sum(times) / len(times)
    value = B_cpu * T / SR / mean
    return dict(value=value, unit="audio-seconds/s", cores=cores, kind="port",
                sample=f"{len(times)} timed phase-2 steps (fwd+bwd, D every 4th) of v2 B={B_cpu}x{T} fp32 on "
                       f"{cores} host threads via oracle/rave_oracle.py (torch CPU)",
                ms_per_step=mean * 1e3, steps=len(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args, max(1, min(args.steps, 4)), 1, 150.0)
    line = {
        "impl": "reference", "metric": "audio-seconds/s (v2 train step fwd+bwd, 48 kHz)",
        "value": r["value"], "unit": "audio-seconds/s", "n_gpus": args.gpus, "steps": r["steps"],
        "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config} phase-2 train step, B=1x65536 sample on host cores (reference "
                               "arithmetic: oracle port; the reference itself needs gin/cached_conv/"
                               "pytorch_lightning which are not installable here)"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "audio-seconds/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def dominant_kernel_roofline(torch, pk):
    """Time the dominant kernel alone: the fp32 implicit-GEMM conv on the hottest block shape
    (DilatedUnit conv3, C=96, L=4096, B=32, d=1).  Inputs (50 MB) + outputs (50 MB)
    exceed nothing but are re-written between launches by rotating over 4 buffer sets > L2 (126 MB)."""
    from rave_b200 import ops
    B, C, L, K = 32, 96, 4096, 3
    xs = [torch.randn(B, C, L, device="cuda") for _ in range(4)]
    w = torch.randn(C, C, K, device="cuda") * 0.05
    ys = [torch.empty(B, C, L, device="cuda") for _ in range(4)]
    for i in range(4):
        ops._gather(xs[i], w, None, None, ys[i], K, 1, 1, 1, C * K, K, 1, 0.2, None)
    torch.cuda.synchronize()
    n = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ops._gather(xs[i % 4], w, None, None, ys[i % 4], K, 1, 1, 1, C * K, K, 1, 0.2, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * B * L * C * C * K
    byts = 4.0 * (2 * B * C * L + C * C * K)
    t_hbm = byts / (pk["hbm_gbs"] * 1e9)
    t_tensor = flops / (pk["bf16_tflops"] * 1e12)
    bound = "hbm" if t_hbm >= t_tensor else "tensor"
    if bound == "hbm":
        ach, peak, unit = byts / (ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
    else:
        ach, peak, unit = flops / (ms * 1e-3) / 1e12, pk["bf16_tflops"], "TFLOP/s"
    return dict(bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak, traffic=None,
                kernel="conv_f32_kernel<0> (act+conv3 of DilatedUnit C=96 L=4096 B=32)",
                ms_per_launch=ms, algorithmic_bytes=byts, algorithmic_flops=flops,
                peak_source=pk["source"] + " burst", note="fp32 CUDA-core parity kernel")


def tc_kernel_roofline(torch, pk):
    """Dominant kernel of the bf16 step: the tcgen05 implicit-GEMM conv, timed alone on the MSD
    384->768 k15 s4 layer at the BASELINE batch (real+fake = 64 x 1024 rows in, 256 rows out).
    4 rotating buffer sets (4 x 125 MB > 126 MB L2) so operands come from HBM."""
    from rave_b200 import ops
    B, Cin, Cout, Lin, K, stride, pad = 64, 384, 768, 1024, 15, 4, 7
    Lout = (Lin + 2 * pad - K) // stride + 1
    xs = [torch.randn(B, Lin, Cin, device="cuda").bfloat16() for _ in range(4)]
    wt = (torch.randn(K, Cout, Cin, device="cuda") * 0.02).bfloat16()
    of = [torch.empty(B, Lout, Cout, device="cuda") for _ in range(4)]
    oa = [torch.empty(B, Lout, Cout, device="cuda", dtype=torch.bfloat16) for _ in range(4)]

    def run(i):
        ops.conv1d_tc(xs[i % 4], wt, None, None, stride, 1, (pad, pad), 1, 0.2, want_f32=False, want_act=False,
                      out_f32=of[i % 4], out_act=oa[i % 4], Lout=Lout)
    for i in range(4):
        run(i)
    torch.cuda.synchronize()
    n = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * B * Lout * Cout * Cin * K
    byts = 2.0 * B * Lin * Cin + B * Lout * Cout * (4 + 2) + 2.0 * K * Cout * Cin
    t_hbm = byts / (pk["hbm_gbs"] * 1e9)
    t_tensor = flops / (pk["bf16_tflops"] * 1e12)
    bound = "hbm" if t_hbm >= t_tensor else "tensor"
    if bound == "hbm":
        ach, peak, unit = byts / (ms * 1e-3) / 1e9, pk["hbm_gbs"], "GB/s"
    else:
        ach, peak, unit = flops / (ms * 1e-3) / 1e12, pk["bf16_tflops"], "TFLOP/s"
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this launch from the committed `ncu --set full` capture
    # (profiles/r1_ncu_conv_tc2_msd384_768.md): 64.7 MB + 35.0 MB, below the algorithmic bytes (outputs partly in L2)
    return dict(bound=bound, achieved=ach, peak=peak, unit=unit, frac=ach / peak, traffic=99.7e6,
                traffic_source="profiles/r1_ncu_conv_tc2_msd384_768.md (ncu --set full, one launch)",
                kernel="conv_tc2_kernel<256,64> (MSD conv 384->768 k15 s4, B=64, Lin=1024)",
                ms_per_launch=ms, algorithmic_bytes=byts, algorithmic_flops=flops,
                peak_source=pk["source"] + " burst (cuBLAS bf16)")


def forward_roofline(torch, model, x_dev, pk, prec):
    """North-star sub-metric: PQMF + encoder + generator FORWARD (v2, B=32x65536), against the block-fused
    algorithmic work of SURVEY.md 8d (306.4 GFLOP, 1.690 GB per 32x65536 batch) and the measured peaks:
    t_min = max(bytes / HBM, flops / tensor) as a single-kernel-equivalent bound."""
    B = x_dev.shape[0]
    flops = 306.4e9 * B / 32
    byts = 1.690e9 * B / 32
    with torch.no_grad():
        for _ in range(3):
            model(x_dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            model(x_dev)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t_hbm = byts / (pk["hbm_gbs"] * 1e9) * 1e3
    t_tc = flops / (pk["bf16_tflops"] * 1e12) * 1e3
    t_min = max(t_hbm, t_tc)
    return dict(ms=ms, audio_seconds_per_s=B * T / SR / (ms * 1e-3), algorithmic_gflop=flops / 1e9,
                algorithmic_gb=byts / 1e9, achieved_tflops=flops / (ms * 1e-3) / 1e12,
                achieved_gbs=byts / (ms * 1e-3) / 1e9, t_min_ms=t_min, frac_of_roofline=t_min / ms,
                note="eager launches, no CUDA graph; byte count is the fp32 block-fused formula of SURVEY 8d")


def stock_cudnn_step(torch, args, B):
    """The reference's own GPU execution path for this step: ATen -> cuDNN convolutions with TF32 enabled
    (scripts/train.py:135-136), fp32 tensors, eager autograd.  Executed through the oracle restatement (the
    reference modules need gin / cached_conv / pytorch_lightning); a reported baseline, never the product."""
    from oracle import rave_oracle as O
    from rave_b200 import configs
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.set_float32_matmul_precision("high")
    torch.manual_seed(0)
    m = configs.build_rave(args.config, sampling_rate=SR)
    sd = {k: v.detach().clone().cuda() for k, v in m.state_dict().items()}
    del m
    cfg = O.ArchConfig() if args.config == "v2" else O.v2_small_config()
    x = synthetic_batch(B).cuda()
    import numpy as np
    eps = torch.randn(B, cfg.latent_size, T // (16 * int(np.prod(cfg.ratios))), device="cuda")
    for i in range(3):
        O.train_step_cpu(x, sd, cfg, eps, dis_step=(i % 4 == 0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 8
    e0.record()
    for i in range(n):
        O.train_step_cpu(x, sd, cfg, eps, dis_step=(i % 4 == 0))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return dict(ms_per_step=ms, audio_seconds_per_s=B * T / SR / (ms * 1e-3),
                note="oracle restatement on CUDA tensors: ATen/cuDNN TF32 convs, eager autograd, fwd+bwd without "
                     "optimiser; same D-every-4th schedule")


def run_ours(args):
    import torch
    import torch.distributed as dist
    import rave_b200
    from rave_b200 import _lib, configs, ddp
    prec = "bf16" if args.precision in ("auto", "bf16") else "fp32"
    rave_b200.set_precision(prec)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pk = peaks()

    torch.manual_seed(0)
    model = configs.build_rave(args.config, sampling_rate=SR).cuda().train()
    model.warmed_up = True          # phase 2: discriminator in the loop (BASELINE config 3)
    ddp.broadcast_module(model)
    reducer = ddp.GradientAllReducer(async_op=args.no_graphs) if world > 1 else None
    B = args.batch
    x_host = synthetic_batch(B, seed=1234 + rank).pin_memory()
    x_dev = x_host.cuda()
    # second resident batch so consecutive steps do not hit identical cache lines
    x_dev2 = synthetic_batch(B, seed=4321 + rank).cuda()

    trainer = None
    graph_note = "eager launches"
    if not args.no_graphs:
        try:
            from rave_b200.graphs import GraphedTrainer
            trainer = GraphedTrainer(model, x_dev, grad_hook=reducer)
            graph_note = "whole-step CUDA graphs (one for the G-step, one for the D-step)"
        except Exception as e:          # report, and measure the eager path instead
            graph_note = f"eager launches (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            trainer = None
            model._optimizers = None
            torch.cuda.synchronize()

    def step(i, x):
        if trainer is not None:
            return trainer.step(x, i)
        return model.training_step(x, i, grad_hook=reducer)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i, x_dev if i % 2 == 0 else x_dev2)
        model.on_train_batch_end()
    barrier()

    # ---- device-resident timed region -------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i, x_dev if i % 2 == 0 else x_dev2)
        model.on_train_batch_end()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - n0
    if trainer is not None:      # replayed launches do not pass through the library's counter
        launches += sum(trainer.launches[model.is_discriminator_step(i)] for i in range(args.steps))
    clocks = sampler.stop() if sampler else None

    # ---- end to end: pinned host input -> H2D -> step -> D2H of the loss --------------------
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    # Every step copies its input from pinned host memory and its result back to pinned host memory; the host waits for
    # the result of step i-1 while step i runs (asynchronous logging, as a training loop would do it) -- a blocking
    # read after every step would expose the host cost of launching the (multi-stream) step graph.
    d2h_bytes = 4
    pinned = torch.empty(args.steps, dtype=torch.float32).pin_memory()
    evs, host_vals = [], []
    t0.record()
    for i in range(args.steps):
        # graphs: the trainer copies the pinned batch straight into the graph's static input (one H2D copy)
        xb = x_host if trainer is not None else x_host.cuda(non_blocking=True)
        logs = step(i, xb)
        model.on_train_batch_end()
        key = "loss_dis" if model.is_discriminator_step(i) else "fullband_spectral_distance"
        pinned[i:i + 1].copy_(logs[key].detach().float().reshape(1), non_blocking=True)   # D2H of the step's result
        ev = torch.cuda.Event()
        ev.record()
        evs.append(ev)
        if i >= 1:
            evs[i - 1].synchronize()
            host_vals.append(float(pinned[i - 1]))
    evs[-1].synchronize()
    host_vals.append(float(pinned[args.steps - 1]))
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    times = torch.tensor([ms, ms_e2e], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e = times.tolist()
    if world > 1 and rank != 0:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    audio_s = world * B * T / SR
    value = audio_s * args.steps / (ms * 1e-3)
    value_e2e = audio_s * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        roof = tc_kernel_roofline(torch, pk) if prec == "bf16" else dominant_kernel_roofline(torch, pk)
        fwd = forward_roofline(torch, model, x_dev, pk, prec)
        stock = None
        if args.cudnn_baseline and world == 1:
            del model, trainer
            torch.cuda.empty_cache()
            try:
                stock = stock_cudnn_step(torch, args, B)
            except Exception as e:
                stock = {"error": f"{type(e).__name__}: {str(e)[:160]}"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_reference_run(args, 2, 1, args.cpu_seconds)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": "audio-seconds/s (v2 train step fwd+bwd, 48 kHz)",
            "value": value, "unit": "audio-seconds/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if prec == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"{args.config} phase-2 training step (PQMF+enc+gen+MPD/MSD disc, fwd+bwd+Adam; "
                                   f"1 D-step per 4), per-GPU batch {B}x{T} @48kHz",
                       "global_batch": world * B, "samples": T, "parallelism": f"dp{world}",
                       "precision": ("bf16 operands / fp32 accumulate (tcgen05 engine); PQMF + losses fp32"
                                     if prec == "bf16" else "fp32 parity kernels (CUDA-core FMA)"),
                       "launch": graph_note,
                       "l2_policy": "working set per step (>10 GB of activations) exceeds the 126 MB L2; "
                                    "two alternating input batches"},
            "roofline": roof, "forward_pqmf_enc_gen": fwd, "stock_cudnn_tf32": stock, "cpu_baseline": cpu,
            "e2e": {"value": value_e2e, "unit": "audio-seconds/s", "h2d_bytes_per_step": B * T * 4,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        # The captured graphs hold NCCL kernels; tearing the communicator down under them can block
        # (observed: both ranks stuck in destroy_process_group after the result line).  Synchronise, flush, and
        # leave without running the communicator's destructor.  No collective after the timing reduction: ranks
        # other than 0 are done there and leave on their own (see below), rank 0 finishes its single-GPU extras alone.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
