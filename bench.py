#!/usr/bin/env python
"""bench.py — AutoVC Generator training throughput (utterance-crops/sec) on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)
    python bench.py --impl torch-gpu --steps K                       # comparison: the same nn.Module tree on stock PyTorch/cuDNN (GPU)

Workload (BASELINE.json configs[1]): Generator(dim_neck=16, dim_emb=256, dim_pre=512, freq=16),
synthetic 80-bin mel crops, batch 256 per GPU, len_crop 128; one step = solver_encoder.py:228-300
(two Generator calls, three losses, zero_grad, backward, Adam.step, gradient all-reduce when N>1).
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for the definition of every key.

Secondary legs (not the driver's default): --workload frontend | convert (BASELINE.json configs[4], 4096 x 10 s utterances),
--workload loader | dvector | wav (the rows SURVEY 8(f) marks "next"); --n-bins 513 (configs[3]); --dim-neck 32 --freq 32 --batch 128
--len-crop 256 (configs[2]); --precision fp32 (3xTF32 on the tensor cores) | fp32_simt (CUDA cores) | tf32 | half.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

# forward MACs per frame of the module definitions; train FLOPs = 3 * 2 * MAC
MAC_PER_FRAME = {(16, 80): 31_784_960, (32, 80): 32_030_720, (16, 513): 36_662_272, (32, 513): 36_908_032}


def synth_batch(B, T, n_bins, dim_emb, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, T, n_bins, generator=g)
    e = F.normalize(torch.randn(B, dim_emb, generator=g), dim=-1) * 0.8
    return x, e


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"),
                "hbm": d.get("hbm_gbs"), "source": "MEASURED_PEAKS.json"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's PyTorch CPU path
# ----------------------------------------------------------------------------------------
def cpu_reference_throughput(dim_neck, freq, T, sample_B, steps, warmup, threads=None, n_bins=80, full_B=0):
    """Oracle port of the reference's PyTorch CPU training step on ALL host cores (torchrun exports OMP_NUM_THREADS=1 to
    its workers: the thread count is therefore set explicitly).  `steps` timed steps on a `sample_B`-crop slice of the batch;
    with `full_B` one additional step at the full per-GPU batch is timed on its own (its crops/s is reported next to the
    sample's, so the two batch sizes can be compared)."""
    from oracle import generator_ref as gref
    torch.set_num_threads(threads or os.cpu_count() or 1)
    torch.manual_seed(0)
    G = gref.build_reference_like_module(dim_neck, 256, 512, freq, n_bins=n_bins).train()
    opt = torch.optim.Adam(G.parameters(), 1e-4)
    x, e = synth_batch(sample_B, T, n_bins, 256, 1234)
    for _ in range(warmup):
        gref.module_train_step(G, opt, x, e)
    t0 = time.perf_counter()
    for _ in range(steps):
        gref.module_train_step(G, opt, x, e)
    dt = (time.perf_counter() - t0) / steps
    full = None
    if full_B and full_B != sample_B:
        xf, ef = synth_batch(full_B, T, n_bins, 256, 1234)
        t0 = time.perf_counter()
        gref.module_train_step(G, opt, xf, ef)
        dtf = time.perf_counter() - t0
        full = {"batch": full_B, "crops_per_s": full_B / dtf, "s_per_step": dtf, "steps": 1}
    return sample_B / dt, dt, torch.get_num_threads(), full


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = args.cpu_sample_batch
    val, dt, threads, full = cpu_reference_throughput(args.dim_neck, args.freq, args.len_crop, sample_B, args.steps, args.warmup,
                                                      n_bins=args.n_bins, full_B=args.batch)
    line = {
        "impl": "reference", "metric": "AutoVC train utterance-crops/sec", "value": val, "unit": "crops/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": val, "unit": "crops/s", "cores": threads, "kind": "port",
                         "sample": f"each timed step is a {sample_B}-crop x {args.len_crop}-frame slice of the {args.batch}-crop batch "
                                   f"({args.warmup} warm-up + {args.steps} timed steps), torch {torch.__version__} CPU (oneDNN), fp32, {threads} threads, "
                                   f"through the oracle port of model_vc_mel.py + solver_encoder.py:228-300; full_batch = one step at the whole batch",
                         "full_batch": full},
        "e2e": {"value": val, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_torch_gpu(args):
    """Comparison arm asked for by the round-1 review (SURVEY 2.3's bar): the SAME nn.Module tree the CPU baseline uses
    (torch.nn Conv1d / BatchNorm1d / LSTM / Linear), moved to the GPU, i.e. cuDNN / cuBLAS kernels through stock PyTorch, in
    three settings: fp32 (TF32 off), TF32 on, bf16 autocast.  Single GPU; not the product path."""
    from oracle import generator_ref as gref
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, T = args.batch, args.len_crop
    x, e = synth_batch(B, T, args.n_bins, 256, 1234)
    x, e = x.to(dev), e.to(dev)
    res = {}
    for mode in ("fp32", "tf32", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
        torch.backends.cudnn.allow_tf32 = mode != "fp32"
        torch.manual_seed(0)
        G = gref.build_reference_like_module(args.dim_neck, 256, 512, args.freq, n_bins=args.n_bins).to(dev).train()
        opt = torch.optim.Adam(G.parameters(), 1e-4, fused=True)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16_autocast")):
                x_identic, x_identic_psnt, code_real = G(x, e, e)
                code_reconst = G(x_identic_psnt, e, None)
            l = F.mse_loss(x, x_identic.float().squeeze(1)) + F.mse_loss(x, x_identic_psnt.float().squeeze(1)) + \
                F.l1_loss(code_real.float(), code_reconst.float())
            opt.zero_grad(set_to_none=True)
            l.backward()
            opt.step()
        for _ in range(max(3, args.warmup)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        res[mode] = {"ms_per_step": ms, "crops_per_s": B / (ms * 1e-3)}
    best = max(res.values(), key=lambda r: r["crops_per_s"])
    line = {"impl": "torch-gpu", "metric": "AutoVC train utterance-crops/sec", "value": best["crops_per_s"], "unit": "crops/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32 / tf32 / bf16 autocast (best reported as value)", "data": "synthetic",
            "config": workload_config(args), "modes": res,
            "note": f"stock PyTorch {torch.__version__} (cuDNN conv/LSTM, cuBLAS, fused Adam) on the same module tree, inputs resident in HBM"}
    print(json.dumps(line), flush=True)


def workload_config(args):
    """The same dict in every arm (the driver compares them): the precision of the arm is the line's `dtype`."""
    return {"workload": f"AutoVC mel Generator train step (solver_encoder.py:228-300), dim_neck={args.dim_neck} dim_emb=256 dim_pre=512 "
                        f"freq={args.freq}, synthetic {args.n_bins}-bin {'mel' if args.n_bins == 80 else 'linear spectrogram'}, batch {args.batch} per GPU, len_crop {args.len_crop}",
            "global_batch": args.batch * args.gpus, "len_crop": args.len_crop, "parallelism": f"dp{args.gpus}",
            "l2": "per-step working set (>5 GB of activations) far exceeds the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def kernel_flops(name, a):
    """Algorithmic FLOPs of one C-ABI call from its scalar arguments (see include/autovc_b200.h)."""
    if name == "avc_gemm_nt_taps":
        nB, T, N, K, taps = a[6], a[7], a[8], a[9], a[10]
        return 2.0 * nB * T * N * K * taps
    if name == "avc_gemm_tn_taps":
        nB, T, N, K, taps = a[5], a[6], a[7], a[8], a[9]
        return 2.0 * nB * T * N * K * taps
    if name in ("avc_gemm_nt_taps_h", "avc_gemm_tn_taps_h"):
        nB, T, N, K, taps = a[7], a[8], a[9], a[10], a[11]
        return 2.0 * nB * T * N * K * taps
    if name == "avc_gemm_nt_taps_hw":
        nB, T, N, K, taps = a[9], a[10], a[11], a[12], a[13]
        return 2.0 * nB * T * N * K * taps
    if name == "avc_lstm_seq_fwd_h":
        nB, T, H = a[10], a[11], a[12]
        return 2.0 * nB * T * 4 * H * H
    if name == "avc_lstm_seq_bwd_h":
        nB, T, H = a[8], a[9], a[10]
        return 2.0 * nB * T * 4 * H * H
    if name == "avc_lstm_seq_fwd":
        nB, T, H = a[6], a[7], a[8]
        return 2.0 * nB * T * 4 * H * H
    if name == "avc_lstm_seq_bwd":
        nB, T, H = a[7], a[8], a[9]
        return 2.0 * nB * T * 4 * H * H
    return 0.0


def family_traffic(family, args):
    """dram__bytes_read + dram__bytes_write per launch (per C-ABI call) of the dominant family at config 2, from committed ncu
    counters: the persistent recurrences from app-range replay (profiles/r02c_traffic.json: one launch each at B=256, T=128, H=512
    and H=1024; a step launches the family three times -- lstm1 at H=512, the two lstm2 layers at H=1024 -- the average over
    those), the GEMM / BatchNorm families from the ncu launch list of one bench step taken with the dram__bytes metrics
    (profiles/r02c_family_traffic.json).  None for other families / shapes."""
    if (args.batch, args.len_crop, args.n_bins) != (256, 128, 80):
        return None
    which = {"avc_lstm_seq_fwd_h": "fwd", "avc_lstm_seq_bwd_h": "bwd"}.get(family)
    if which is not None:
        path = os.path.join(ROOT, "profiles", "r02c_traffic.json")
        if not os.path.exists(path):
            return None
        t = json.load(open(path))["per_launch"]
        return (t[f"{which}_h512"]["dram_bytes"] + 2 * t[f"{which}_h1024"]["dram_bytes"]) / 3.0
    path = os.path.join(ROOT, "profiles", "r02c_family_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path))["per_call"].get(family)


def run_ours(args):
    import torch.distributed as dist
    import autovc_b200
    from autovc_b200 import _lib, solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        solver.nccl_env_defaults()      # cap NCCL's CTAs so that its kernels co-reside with the 128-CTA recurrence kernels
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(0)
    if args.n_bins == 513:      # BASELINE.json configs[3]: the model_vc_stft variant (its .model)
        G = autovc_b200.GeneratorSTFT(args.dim_neck, 256, 512, args.freq, precision=args.precision).model.to(dev).train()
    else:
        G = autovc_b200.Generator(args.dim_neck, 256, 512, args.freq, precision=args.precision).to(dev).train()
    opt = autovc_b200.FusedAdam(G.parameters(), 1e-4)      # solver_encoder.py:130's Adam, one-launch step
    reducer = None
    if world > 1:
        solver.broadcast_parameters(G)
        reducer = solver.GradBucketReducer(G.parameters(), bucket_mb=args.bucket_mb, nccl_max_ctas=args.nccl_max_ctas)
    B, T = args.batch, args.len_crop
    x_host, e_host = synth_batch(B, T, args.n_bins, 256, 1234 + rank)
    x_pin, e_pin = x_host.pin_memory(), e_host.pin_memory()
    x_dev, e_dev = x_host.to(dev), e_host.to(dev)

    def step_resident():
        return solver.train_step(G, opt, x_dev, e_dev, reducer=reducer, sync_losses=False)

    prefetch = solver.HostBatchPrefetcher(dev)
    prefetch.put(x_pin, e_pin)
    pending = [None]

    def step_e2e():
        # every step copies one batch from pinned host memory (the copy of the NEXT step's batch is enqueued on a side
        # stream before this step's kernels, so K timed steps contain K host->device copies) and reads the losses back
        xd, ed = prefetch.get()
        prefetch.put(x_pin, e_pin)
        # D2H of the losses every step (like :315-317), into pinned memory; the host reads step n's values while step n+1
        # is already queued, so the GPU never waits for the host between steps
        res = solver.train_step(G, opt, xd, ed, reducer=reducer, sync_losses="async")
        prev, pending[0] = pending[0], res["losses"]
        return prev.values() if prev is not None else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # the sampler starts BEFORE the warm-up: nvidia-smi's NVML initialisation briefly stalls every GPU of the box, which must
    # not fall into the timed region (it cost one of two timed loops ~35 ms on a 4-GPU box)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    if sampler:
        torch.cuda.synchronize()
        sampler.rows.clear()            # keep only the samples taken during the timed region
    n0 = autovc_b200.launch_count()
    mem0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    ms = timed(step_resident, args.steps)
    if os.environ.get("AVC_BENCH_MEMSTATS"):     # cudaMalloc calls inside the timed region stall the step: should be 0
        st = torch.cuda.memory_stats(dev)
        print(f"[memstats] cudaMalloc calls in the timed region: {st.get('num_device_alloc', 0) - mem0}, reserved "
              f"{st.get('reserved_bytes.all.current', 0) / 2**30:.1f} GiB, retries {st.get('num_alloc_retries', 0)}", file=sys.stderr, flush=True)
    launches = (autovc_b200.launch_count() - n0) // args.steps
    clocks = sampler.stop() if sampler else None
    for _ in range(3):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # per-kernel-family device time over instrumented steps (CUDA events on the launching stream)
    prof_steps = max(1, min(3, args.steps))
    _lib.enable_timing(True)
    for _ in range(prof_steps):
        step_resident()
    torch.cuda.synchronize()
    records = _lib.collect_timing()
    _lib.enable_timing(False)
    fam = {}
    for name, scalars, ms_k in records:
        f = fam.setdefault(name, {"ms": 0.0, "n": 0, "flops": 0.0})
        f["ms"] += ms_k
        f["n"] += 1
        f["flops"] += kernel_flops(name, scalars)
    tot_ms = sum(f["ms"] for f in fam.values()) or 1.0
    top = max(fam.items(), key=lambda kv: kv[1]["ms"])
    peaks = measured_peaks()
    step_flops = 6.0 * MAC_PER_FRAME.get((args.dim_neck, args.n_bins), MAC_PER_FRAME[(16, 80)]) * B * T
    top_name, top_f = top
    achieved = top_f["flops"] / (top_f["ms"] * 1e-3) / 1e12 if top_f["ms"] > 0 else 0.0
    roofline = {
        "bound": "tensor", "kernel": top_name, "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_sustained"], "traffic": family_traffic(top_name, args),
        "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
        "launches_per_step": top_f["n"] / prof_steps, "avg_launch_ms": top_f["ms"] / max(1, top_f["n"]),
        "share_of_step_kernel_time": top_f["ms"] / tot_ms,
        "step": {"flops": step_flops, "achieved_tflops": step_flops / (ms * 1e-3) / 1e12,
                 "frac": step_flops / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"]},
        "families": {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["n"] / prof_steps,
                         "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None}
                     for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])[:(64 if os.environ.get("AVC_BENCH_ALL_FAMILIES") else 8)]},
    }

    # ---- sustained leg: a seconds-long run with its own clock trace (the timed region above is K steps = a fraction of a
    # second at boost clock; this shows whether the rate holds once the power limit has had time to act)
    sustained = None
    if args.sustained > 0:
        samp2 = ClockSampler(local) if rank == 0 else None
        if samp2:
            samp2.start()
            time.sleep(0.5)             # NVML initialisation of the sampler stays out of the region
            samp2.rows.clear()
        ms_s = timed(step_resident, args.sustained)
        ck = samp2.stop() if samp2 else None
        trace = [float(r[0]) for r in samp2.rows if r and r[0].replace(".", "", 1).isdigit()] if samp2 else []
        sustained = {"steps": args.sustained, "ms_per_step": ms_s, "value": B * world / (ms_s * 1e-3), "unit": "crops/s", "clocks": ck,
                     "sm_mhz_trace_200ms": trace[:200]}

    # ---- the other precision modes of the same module, same process, same inputs (N=1 only: keeps the scaling runs short)
    modes = {args.precision: {"ms_per_step": ms, "value": B * world / (ms * 1e-3), "steps": args.steps}}
    if world == 1 and args.modes:
        for mode in [m for m in ("tf32", "fp32", "fp32_simt", "half") if m != args.precision]:
            G.set_precision(mode)
            slow = mode == "fp32_simt"          # CUDA-core cross-check of the parity mode: ~0.5 s per step
            k = args.steps if not slow else max(2, min(args.steps, 3))
            for _ in range(3 if not slow else 1):
                step_resident()
            ms_m = timed(step_resident, k)
            modes[mode] = {"ms_per_step": ms_m, "value": B / (ms_m * 1e-3), "steps": k}
        G.set_precision(args.precision)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            val, dt, threads, full = cpu_reference_throughput(args.dim_neck, args.freq, T, args.cpu_sample_batch, 2, 1, n_bins=args.n_bins, full_B=B)
            cpu = {"value": val, "unit": "crops/s", "cores": threads, "kind": "port",
                   "sample": f"{args.cpu_sample_batch} crops x {T} frames per step (a slice of the {B}-crop batch), 1 warm-up + 2 timed steps, "
                             f"oracle port of the reference's PyTorch CPU path (fp32, {threads} threads); full_batch = one step at the whole batch",
                   "full_batch": full}
        line = {
            "metric": "AutoVC train utterance-crops/sec", "value": B * world / (ms * 1e-3), "unit": "crops/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "crops/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": (x_pin.numel() + e_pin.numel()) * 4, "d2h_bytes_per_step": 16},
            "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "modes": modes, "sustained": sustained,
        }
        if reducer is not None:
            line["config"]["nccl_cta_cap"] = reducer.cta_cap
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_frontend_or_convert(args):
    """Secondary legs (not the driver's default): `--workload frontend` = make_spect log-mel front-end alone
    (HBM roofline, SURVEY 8(d): 1 480 320 algorithmic bytes per 10 s utterance); `--workload convert` =
    BASELINE.json configs[4]: front-end -> pad to x32 -> eval forward with swapped embeddings, mel frames/sec."""
    import numpy as np
    import autovc_b200
    from autovc_b200.conversion import convert, padded_frames
    from autovc_b200.make_spect import Spect
    from oracle import make_spect_ref as fref
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and not (world == 1 and args.gpus > 1):
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        # replicas only: the utterances are split across the ranks, no data-path collective; the process group
        # serves the barrier and the max-over-ranks of the device-timed region
        dist.init_process_group("nccl", device_id=dev)
    n_total, L = args.utterances, 160000
    n = n_total // world
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    env = 0.55 + 0.45 * torch.sin(torch.linspace(0, 60, L, device=dev))[None, :]
    wav = (0.1 * torch.randn(n, L, generator=g, device=dev) * env).clamp(-1.0, 1.0 - 2 ** -15)
    dither = torch.rand(n, L, generator=g, device=dev)
    sp = Spect()
    F_ = 1 + L // 256
    peaks = measured_peaks()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    if args.workload == "frontend":
        ms = timed(lambda: sp.logmel(wav, dither, None, max_frames=padded_frames(L)), args.steps, args.warmup)
        alg_bytes = n * (4 * L + 4 * L + 4 * 80 * F_)
        gbs = alg_bytes / (ms * 1e-3) / 1e9          # per GPU (the roofline is a per-device figure)
        if rank != 0:
            dist.destroy_process_group()
            return
        # CPU baseline: oracle restatement on a bounded sample, single process like the reference
        k = min(8, n)
        wc, dc = wav[:k].cpu().numpy(), dither[:k].double().cpu().numpy()
        t0 = time.perf_counter()
        for i in range(k):
            fref.logmel_from_wav(wc[i], dc[i])
        cpu_fps = k * F_ / (time.perf_counter() - t0)
        line = {"metric": "log-mel front-end mel frames/sec", "value": n * world * F_ / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64 IIR + f32 FFT", "data": "synthetic",
                "config": {"workload": f"make_spect front-end, {n_total} synthetic 10 s 16 kHz waveforms, {n} per GPU (inputs {2*n*L*4/1e9:.2f} GB per GPU > L2)",
                           "parallelism": f"replicas x{world}"},
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                             "traffic": None, "peak_source": peaks["source"] + " hbm_gbs",
                             "algorithmic_bytes_per_utterance": 4 * L + 4 * L + 4 * 80 * F_},
                "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": 1, "kind": "port",
                                 "sample": f"{k} of the {n} utterances through oracle/make_spect_ref.py (scipy filtfilt + numpy rfft)"}}
        print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    torch.manual_seed(0)
    G = autovc_b200.Generator(32, 256, 512, 32, precision=args.precision).to(dev)
    e = F.normalize(torch.randn(n, 256, generator=torch.Generator().manual_seed(5)), dim=-1).to(dev) * 0.8
    et = e.roll(1, 0).contiguous()
    ms = timed(lambda: convert(G, sp, wav, dither, None, e, et, chunk=args.chunk, streams=args.streams), args.steps, max(1, args.warmup - 2))
    flops = 2.0 * 28_385_280 * n * padded_frames(L)      # per GPU; one encoder pass in conversion (28 385 280 MAC/frame)
    if rank != 0:
        dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # the reference's CPU path on a bounded sample: make_spect restatement (scipy/numpy, single process like the reference)
        # + the reference-like nn.Module in eval mode on all host threads (conversion.py:40-47,:91-92)
        from oracle import generator_ref as gref
        k = 4
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        torch.manual_seed(0)
        Gc = gref.build_reference_like_module(32, 256, 512, 32).eval()
        wc, dc = wav[:k].cpu().numpy(), dither[:k].double().cpu().numpy()
        ec = e[:k].cpu()
        t0 = time.perf_counter()
        mels = [torch.from_numpy(np.pad(fref.logmel_from_wav(wc[i], dc[i]), ((0, padded_frames(L) - F_), (0, 0)))) for i in range(k)]
        t1 = time.perf_counter()
        with torch.no_grad():
            for i in range(k):
                Gc(mels[i][None], ec[i:i + 1], ec.roll(1, 0)[i:i + 1])
        t2 = time.perf_counter()
        cpu = {"value": k * F_ / (t2 - t0), "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"{k} of the {n_total} utterances: oracle make_spect restatement (1 thread, {t1 - t0:.2f} s) + reference-like module eval "
                         f"forward at 640 frames ({threads} threads, {t2 - t1:.2f} s)"}
    line = {"metric": "conversion mel frames/sec", "value": n * world * F_ / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"waveform -> log-mel -> pad to x32 -> Generator(32,256,512,32).eval() forward, {n_total} x 10 s utterances "
                                   f"({n} per GPU), chunks of {args.chunk} over {args.streams} streams", "parallelism": f"replicas x{world}"},
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"], "traffic": None, "per": "GPU"},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_widened(args):
    """Secondary legs for the rows SURVEY 8(f) marks "next": `--workload loader` (crop loader, HBM roofline: 2*4*T*n_bins bytes per
    crop) and `--workload dvector` (speaker encoder forward on 128-frame crops, make_metadata.py:42 shapes)."""
    import numpy as np
    import autovc_b200
    from autovc_b200 import data_loader
    from autovc_b200.model_bl import D_VECTOR
    from oracle import data_loader_ref as lref
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peaks = measured_peaks()
    B, T = args.batch, args.len_crop

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    if args.workload == "loader":
        rs = np.random.RandomState(3)
        corpus = [["p%03d" % s, rs.randn(256).astype(np.float32)] + [rs.rand(int(f), 80).astype(np.float32) for f in rs.randint(64, 900, size=24)]
                  for s in range(512)]                                            # 512 speakers x 24 utterances, ~1.9 GB
        ds = data_loader.Utterances(corpus=corpus, len_crop=T)
        idx = rs.randint(0, len(ds), size=B).tolist()
        sel = ds.draw(idx, rs)
        ms_k = timed(lambda: ds.gather(sel), args.steps * 20, 10)                 # kernel + the 3 KB index upload
        t0 = time.perf_counter()
        for _ in range(50):
            ds.draw(idx, rs)
        ms_draw = (time.perf_counter() - t0) / 50 * 1e3                           # host-side numpy draws (reference order)
        k = 16
        t0 = time.perf_counter()
        for _ in range(5):
            lref.get_batch(corpus, idx[:k], T, rs)
        cpu = k * 5 / (time.perf_counter() - t0)
        alg = B * (2 * 4 * T * 80 + 2 * 4 * 256)
        gbs = alg / (ms_k * 1e-3) / 1e9
        line = {"metric": "crop-loader utterance-crops/sec", "value": B / (ms_k * 1e-3), "unit": "crops/s", "n_gpus": 1, "steps": args.steps * 20,
                "warmup": 10, "ms_per_step": ms_k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (copy)",
                "data": "synthetic", "config": {"workload": f"data_loader batch: {B} crops x {T} frames x 80 bins from a 512-speaker corpus resident in HBM",
                                                "host_draw_ms_per_batch": ms_draw},
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"], "traffic": None,
                             "note": "a 10.5 MB batch is launch-latency sized: one launch, ~2 us of data at the HBM rate"},
                "cpu_baseline": {"value": cpu, "unit": "crops/s", "cores": 1, "kind": "port",
                                 "sample": f"{k}-crop batches through oracle/data_loader_ref.py (numpy restatement of Utterances.__getitem__)"}}
        print(json.dumps(line), flush=True)
        return
    if args.workload == "wav":
        # GeneratorWav training step (model_vc_wav.py / solver_encoder.py:264-300) on waveform crops of (T-1)*256 + 1024 samples
        from autovc_b200 import solver
        L = (T - 1) * 256 + 1024
        torch.manual_seed(0)
        G = autovc_b200.GeneratorWav(args.dim_neck, 256, 512, args.freq, args.depth, precision=args.precision).to(dev).train()
        opt = autovc_b200.FusedAdam(G.parameters(), 1e-4)
        gw = torch.Generator().manual_seed(1234)
        xw = (0.1 * torch.randn(B, L, 1, generator=gw)).clamp(-1, 1).to(dev)          # synthetic waveform crops (B, L, 1)
        ew = (F.normalize(torch.randn(B, 256, generator=gw), dim=-1) * 0.8).to(dev)
        ms = timed(lambda: solver.train_step_wav(G, opt, xw, ew), args.steps, args.warmup)
        # dense MACs per frame: the mel model's encoder (twice) / lstm1 / decoder convs / lstm2 with 512-channel ends, plus the
        # filterbanks (512*1024 each way, analysis twice) and the k=3 layers (depth * 3*512*512, analysis side twice)
        enc = (512 + 256) * 512 * 5 + 2 * 512 * 512 * 5 + 2 * (512 * 4 * args.dim_neck + args.dim_neck * 4 * args.dim_neck) \
            + 2 * (2 * args.dim_neck * 4 * args.dim_neck + args.dim_neck * 4 * args.dim_neck)
        dec = (2 * args.dim_neck + 256 + 512) * 2048 + 3 * 512 * 512 * 5 + (512 + 1024) * 4096 + 2 * 1024 * 4096 + 1024 * 512
        tas = 3 * 512 * 1024 + args.depth * 3 * 3 * 512 * 512
        flops = 3 * 2.0 * (2 * enc + dec + tas) * B * T
        k = 2
        sd = None
        from oracle import generator_wav_ref as wref          # the CPU baseline leg (the only use of oracle/ here)
        torch.set_num_threads(os.cpu_count() or 1)
        torch.manual_seed(0)
        M = wref.build_wav_module(args.dim_neck, 256, 512, args.freq, args.depth)
        sd = {kk: v.detach().clone() for kk, v in M.state_dict().items()}
        xc, ec = xw[:k].cpu(), ew[:k].cpu()
        t0 = time.perf_counter()
        wref.wav_train_step(sd, xc, ec, args.dim_neck, args.freq)
        cpu = k / (time.perf_counter() - t0)
        line = {"metric": "GeneratorWav train utterance-crops/sec", "value": B / (ms * 1e-3), "unit": "crops/s", "n_gpus": 1, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
                "data": "synthetic",
                "config": {"workload": f"GeneratorWav({args.dim_neck},256,512,{args.freq},depth={args.depth}) train step (solver_encoder.py:264-300), "
                                       f"{B} waveform crops x {L} samples ({T} frames)"},
                "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                             "frac": flops / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"], "traffic": None},
                "cpu_baseline": {"value": cpu, "unit": "crops/s", "cores": os.cpu_count() or 1, "kind": "port",
                                 "sample": f"one forward+backward of {k} crops through oracle/generator_wav_ref.py (explicit LSTM time loop; a lower "
                                           f"bound on the reference's oneDNN path)"}}
        print(json.dumps(line), flush=True)
        return
    torch.manual_seed(0)
    C = D_VECTOR(dim_input=80, dim_cell=768, dim_emb=256, precision=args.precision).eval().to(dev)
    x = torch.rand(B, T, 80, device=dev)
    with torch.no_grad():
        ms = timed(lambda: C(x), args.steps, args.warmup)
    mac = 80 * 3072 + 768 * 3072 + 2 * (768 * 3072 * 2) + 768 * 256 / T        # per frame: 3 layers (input proj + recurrence) + the last-frame linear
    flops = 2.0 * mac * B * T
    line = {"metric": "speaker-encoder utterance-crops/sec", "value": B / (ms * 1e-3), "unit": "crops/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic", "config": {"workload": f"D_VECTOR(80, 768, 256) forward, {B} crops x {T} frames (make_metadata.py:42,:77)"},
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": flops / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"], "traffic": None}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-gpu"])
    ap.add_argument("--sustained", type=int, default=500, help="extra timed leg of this many steps with its own clock trace (0: off)")
    ap.add_argument("--no-modes", dest="modes", action="store_false", help="skip the tf32 / fp32 legs of the `modes` key")
    ap.add_argument("--precision", default=os.environ.get("AUTOVC_B200_PRECISION", "half"), choices=["fp32", "fp32_simt", "tf32", "half"])
    ap.add_argument("--batch", type=int, default=256, help="crops per GPU")
    ap.add_argument("--len-crop", dest="len_crop", type=int, default=128)
    ap.add_argument("--dim-neck", dest="dim_neck", type=int, default=16)
    ap.add_argument("--freq", type=int, default=16)
    ap.add_argument("--n-bins", dest="n_bins", type=int, default=80, choices=[80, 513], help="513 = model_vc_stft variant (configs[3])")
    ap.add_argument("--cpu-sample-batch", dest="cpu_sample_batch", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="train", choices=["train", "frontend", "convert", "loader", "dvector", "wav"])
    ap.add_argument("--bucket-mb", type=float, default=25.0, help="N > 1: gradient bucket size of the all-reduce")
    ap.add_argument("--nccl-max-ctas", type=int, default=16, help="N > 1: CTA cap of the reducer's NCCL communicator")
    ap.add_argument("--chunk", type=int, default=256, help="--workload convert: utterances per Generator call")
    ap.add_argument("--streams", type=int, default=2, help="--workload convert: CUDA streams the chunks alternate over")
    ap.add_argument("--depth", type=int, default=1, help="--workload wav: Conv-TasNet encoder/decoder depth (main.py:65)")
    ap.add_argument("--utterances", type=int, default=4096, help="frontend/convert legs: number of 10 s utterances (BASELINE.json configs[4]: 4096)")
    args = ap.parse_args()
    if args.workload in ("loader", "dvector", "wav"):
        return run_widened(args)
    if args.workload != "train":
        return run_frontend_or_convert(args)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-gpu":
        run_torch_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
