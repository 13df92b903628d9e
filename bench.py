#!/usr/bin/env python
"""bench.py — the rx_tools hot path on B200: I/Q Msamples/s through full_demod() / fix_fft().

A "step" = one pass of the hot path over one batch of synthetic CS16 input.

Workloads (BASELINE.json `configs`):
  fm2b   (default; configs[1]) rx_fm -M wbfm -s 300k -F 9 -r 48k: 2.4 Msps capture -> 300 k (3 x fifth_order
         + droop FIR) -> fast atan -> de-emphasis a=23 -> 48 kHz, 1 GiB CS16 stream, chunk 131072 complex
  fm2a   rx_fm -M wbfm -s 2400000 -r 48000 (no decimation, de-emphasis a=181), 1 GiB
  fm1    rx_fm -M fm -s 1024000 -r 24000 (configs[0] shape, D=1, atan2), 256 MiB
  fm5a   256 NBFM channels, boxcar D=100, lut, 2.4 M complex each (channels sharded over ranks)
  power3 rx_power 1 MHz span, 1024-bin, 32768 hop buffers (1 GiB), one hop
  power4 rx_power 24-1766 MHz, 4096-bin, 871 hops x 36 sweeps (1 GiB), hops sharded over ranks + all-gather

Multi-GPU (torchrun, one rank per GPU): rx_fm streams do not shard (serial carry) -> every rank runs its own
stream ("replicas only", weak scaling, no collective).  rx_power hops shard across ranks with ONE all-gather
of the int64 spectrum rows at the end of the step.

--impl reference times the reference's own C path (oracle/_ref, the unmodified sources compiled in the
authoring container; else the port) on the host cores of this box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHUNK = 262144   # int16 per chunk = 131072 complex (MAXIMUM_BUF_LENGTH, src/rtl_fm.c:80-82)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def fm_params(workload):
    from rx_tools_b200 import fm
    if workload == "fm2b":
        return fm.derive_params(wbfm=1, rate_s=300000, rate_r=48000, use_F=1, comp_fir_size=9).params
    if workload == "fm2a":
        return fm.derive_params(wbfm=1, rate_s=2400000, rate_r=48000).params
    if workload == "fm1":
        return fm.derive_params(rate_s=1024000, rate_r=24000).params
    if workload == "fm5a":
        return fm.FmParams(downsample=100, custom_atan=fm.ATAN_LUT, rate_out=24000)
    raise ValueError(workload)


def fm_input_period(workload, n_complex):
    from rx_tools_b200 import synth
    if workload in ("fm2b", "fm2a"):
        return synth.cfg2_iq(n_complex)
    if workload == "fm1":
        return synth.cfg1_iq(n_complex)
    return synth.cfg5_iq(n_complex, 0)


def fm_out_bytes_per_sample(p):
    d = (1 << p.downsample_passes) if p.downsample_passes else p.downsample
    r = (p.rate_out2 / p.rate_out) if p.rate_out2 > 0 else 1.0
    return 2.0 * r / d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2)
        sm, smax, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = max(smax, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [v for v in sm if v > 0.5 * smax] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def barrier_sync(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v, world, device):
    import torch
    if world == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(v, world, device):
    import torch
    if world == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ------------------------------------------------------------------------------------------ rx_fm arm
def run_fm(args, rank, local, world):
    import torch
    from rx_tools_b200 import fm
    dev = torch.device("cuda", local)
    p = fm_params(args.workload)
    if args.workload == "fm5a":
        n_ch_total, n_per = 256, 2_400_000 - (2_400_000 % 8)
        n_ch = n_ch_total // world if world > 1 else (n_ch_total if args.size_mib == 0 else n_ch_total)
        if args.size_mib:
            n_ch = max(1, (args.size_mib << 20) // (n_per * 4))
    else:
        n_ch = 1
        size_mib = args.size_mib or (256 if args.workload == "fm1" else 1024)
        n_per = (size_mib << 20) // 4
    period = min(n_per, 1 << 24)
    period -= period % (CHUNK // 2)
    n_per = (n_per // period) * period if n_per >= period else n_per
    host_period = fm_input_period(args.workload, period)
    reps = n_per // period
    d_period = torch.from_numpy(host_period).to(dev)
    d_in = d_period.repeat(n_ch * reps).contiguous()
    del d_period
    demod = fm.FmDemod(p, device=local, n_channels=n_ch)
    n_int16 = 2 * n_per
    cap = demod.max_output(n_int16, CHUNK) + 8
    d_out = torch.empty(n_ch * cap, dtype=torch.int16, device=dev)
    stream = torch.cuda.ExternalStream(demod.stream, device=dev)

    def step():
        return demod.process_device(d_in.data_ptr(), n_int16, CHUNK, d_out.data_ptr(), cap, sync=False)

    for _ in range(args.warmup):
        n_pcm = step()
    barrier_sync(world)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    e0.record(stream)
    for _ in range(args.steps):
        n_pcm = step()
    e1.record(stream)
    e1.synchronize()
    barrier_sync(world)
    ms = e0.elapsed_time(e1)
    clocks = sampler.finish()
    ms = max_over_ranks(ms, world, dev)
    stats = demod.stats()
    # dominant-kernel duration, CUDA events recorded around the fused kernel on its own stream
    kms = []
    for _ in range(max(3, min(args.steps, 10))):
        step()
        kms.append(demod.kernel_ms())
    kernel_ms = float(np.mean(kms))
    samples_rank = n_ch * n_per
    total_samples = sum_over_ranks(float(samples_rank), world, dev) * args.steps
    value = total_samples / (ms * 1e-3) / 1e6

    # end to end through the public host API: pinned host buffers, H2D + kernel + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(n_ch * n_int16, dtype=torch.int16).pin_memory()
        h_in.view(n_ch * reps, -1)[:] = torch.from_numpy(host_period)
        h_out = torch.empty(n_ch * cap, dtype=torch.int16).pin_memory()
        import ctypes as C
        from rx_tools_b200 import _lib
        npcm = C.c_size_t(0)
        demod.reset()

        def e2e_step():
            _lib.check(_lib.lib().rxb200_fm_process(demod._h, h_in.data_ptr(), n_int16, CHUNK, h_out.data_ptr(), cap,
                                                    C.byref(npcm), None))
        e2e_step()
        barrier_sync(world)
        t0 = time.perf_counter()
        k2 = max(2, min(args.steps, 5))
        for _ in range(k2):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt, world, dev)
        e2e = {"value": sum_over_ranks(float(samples_rank), world, dev) * k2 / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(n_ch * n_int16 * 2), "d2h_bytes_per_step": int(n_ch * npcm.value * 2),
               "steps": k2}
        del h_in, h_out

    peak, peak_src = peaks()
    bytes_per_sample = 4.0 + fm_out_bytes_per_sample(p)
    achieved = samples_rank * bytes_per_sample / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    res = {
        "metric": "I/Q Msamples/s through full_demod() & fix_fft() at 1/2/4/8 B200 vs host CPU",
        "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if (args.workload == "fm5a" and not args.size_mib) else "weak", "vs_baseline": None,
        "dtype": "int16/int32 (fp64 atan2 on first sample of each chunk)", "data": "synthetic",
        "config": {"workload": {"fm2b": "rx_fm -M wbfm -s 300k -F 9 -r 48k (2.4 Msps capture -> 48 kHz), fused kernel",
                                "fm2a": "rx_fm -M wbfm -s 2400000 -r 48000 (D=1), fused kernel",
                                "fm1": "rx_fm -M fm -s 1024000 -r 24000 (atan2), fused kernel",
                                "fm5a": "256 NBFM channels at 2.4 Msps, boxcar D=100, lut"}[args.workload],
                   "stream_bytes_per_gpu": int(n_ch * n_int16 * 2), "channels_per_gpu": n_ch, "chunk_complex": CHUNK // 2,
                   "input": f"synthetic CS16, {period}-sample seeded period tiled, device-resident for `value`",
                   "l2": f"input ({n_ch * n_int16 * 2 / 2**20:.0f} MiB per step) is larger than the 126 MB L2", "parallelism": f"replicas x{world}" if n_ch == 1 else f"channels sharded x{world}",
                   "segment_len": stats["segment_len"], "warmup_len": stats["warmup_len"],
                   "fixup_segments": stats["fixup_segments"]},
        "gpu_launches": stats["launches"] * args.steps,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "fm_fused_kernel", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_sample": bytes_per_sample},
        "e2e": e2e,
    }
    demod.close()
    return res


# ------------------------------------------------------------------------------------------ rx_power arm
def run_power(args, rank, local, world):
    import torch
    from rx_tools_b200 import power, synth
    dev = torch.device("cuda", local)
    if args.workload == "power3":
        plan = power.plan_range("100M:101M:1k")
        window = power.window_table("hann", 1 << plan.bin_e)
        n_pass = ((args.size_mib or 1024) << 20) // (plan.buf_len * 2)
        wl = "rx_power -f 100M:101M:1k, 1024-bin fix_fft, Hann window table, batched hop buffers"
    else:
        plan = power.plan_range("24M:1766M:1k", 0.285)
        window = power.window_table("hamming", 1 << plan.bin_e)
        n_pass = max(1, ((args.size_mib or 1024) << 20) // (plan.n_hops * plan.buf_len * 2))
        wl = "rx_power -f 24M:1766M:1k -c 28.5% -w hamming, 4096-bin fix_fft, 871 hops batched"
    n_hops = plan.n_hops
    # shard hops contiguously over ranks; pad to equal rows for the all-gather
    from rx_tools_b200 import sharding
    per = sharding.rows_per_rank(n_hops, world)
    hb, he = sharding.unit_range(rank, world, n_hops)
    if n_hops == 1:
        hb, he = 0, 1      # one hop: every rank takes a slice of the passes instead
        my_pass = n_pass // world
    else:
        my_pass = n_pass
    nh = he - hb
    base = synth.power_hops(2, min(nh, 8) or 1, plan.buf_len, seed=777 + rank)
    d_base = torch.from_numpy(base.reshape(-1)).to(dev)
    need = my_pass * max(nh, 1) * plan.buf_len
    d_in = d_base.repeat(-(-need // d_base.numel()))[:need].contiguous()
    sc = power.PowerScanner(plan, window, device=local)
    stream = torch.cuda.ExternalStream(sc.stream, device=dev)
    N = 1 << plan.bin_e
    do_gather = world > 1 and n_hops > 1
    src = _device_view(sc.device_avg_ptr + hb * N * 8, max(nh, 1) * N, dev) if do_gather else None
    gathered = [None]

    def step():
        if nh > 0:
            sc.scanner_device(d_in.data_ptr(), my_pass, hb, he, sync=False)
        if do_gather:
            with torch.cuda.stream(stream):      # ordered after the kernel on the handle's stream
                gathered[0] = sharding.gather_rows(src[: nh * N] if nh > 0 else src[:0], n_hops, N, world)

    for _ in range(args.warmup):
        step()
    barrier_sync(world)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    e1.synchronize()
    barrier_sync(world)
    ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    clocks = sampler.finish()
    kms = []
    for _ in range(max(3, min(args.steps, 10))):
        if nh > 0:
            sc.scanner_device(d_in.data_ptr(), my_pass, hb, he, sync=False)
            kms.append(sc.kernel_ms())
    kernel_ms = float(np.mean(kms)) if kms else float("nan")
    samples_rank = my_pass * nh * (plan.buf_len // 2)
    total = sum_over_ranks(float(samples_rank), world, dev) * args.steps
    value = total / (ms * 1e-3) / 1e6
    e2e = None
    if not args.no_e2e and nh > 0:
        h_in = torch.empty(need, dtype=torch.int16).pin_memory()
        h_in.copy_(d_in.cpu())
        k2 = max(2, min(args.steps, 5))
        sc.reset()
        sc.scanner(h_in.numpy(), my_pass, hb, he)
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(k2):
            sc.scanner(h_in.numpy(), my_pass, hb, he)
            avg, smp = sc.read()
        dt = max_over_ranks(time.perf_counter() - t0, world, dev)
        e2e = {"value": sum_over_ranks(float(samples_rank), world, dev) * k2 / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(need * 2), "d2h_bytes_per_step": int(avg.nbytes), "steps": k2}
    peak, peak_src = peaks()
    achieved = samples_rank * 4.0 / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", f"traffic_{wl}.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    fft8 = (plan.downsample == 1 and plan.buf_len == 16384 and 3 <= plan.bin_e <= 13
            and not os.environ.get("RXB200_POWER_V1"))
    res = {
        "metric": "I/Q Msamples/s through full_demod() & fix_fft() at 1/2/4/8 B200 vs host CPU",
        "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int16 FFT, int64 accumulate", "data": "synthetic",
        "config": {"workload": wl, "hops": n_hops, "passes": n_pass, "bins": N, "buf_len_int16": plan.buf_len,
                   "bytes_per_step_all_gpus": int(n_pass * n_hops * plan.buf_len * 2),
                   "l2": "hop buffers per step (~1 GiB) are larger than the 126 MB L2",
                   "parallelism": f"hops sharded x{world} + one all_gather of int64 rows" if n_hops > 1 else f"passes sharded x{world}"},
        "gpu_launches": args.steps, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "power_fft8_kernel" if fft8 else "power_fft_kernel", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_sample": 4.0},
        "e2e": e2e,
    }
    sc.close()
    return res


def _device_view(ptr, n_int64, dev):
    """torch int64 view over raw device memory owned by librxb200 (the accumulator rows)."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n_int64,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=dev)


# ------------------------------------------------------------------------------------------ CPU legs
def _cpu_worker(job):
    kind_pref, workload, n_complex, repeats = job
    import oracle
    from rx_tools_b200 import fm, power, synth  # host-side derivation only (no GPU use)
    if workload.startswith("fm"):
        p = fm_params(workload)
        op = oracle.FmParams(**p.reference_fields())
        x = fm_input_period(workload, n_complex)
        if kind_pref == "reference" and oracle.have_ref():
            t = oracle.RefFm().time(op, x, CHUNK, repeats)
            return t, n_complex * repeats, "reference"
        t = oracle.port().fm_time(op, x, CHUNK, repeats)
        return t, n_complex * repeats, "port"
    if workload == "power3":
        arg, crop, wname = "100M:101M:1k", 0.0, "hann"
    else:
        arg, crop, wname = "24M:1766M:1k", 0.285, "hamming"
    plan = power.plan_range(arg, crop)
    n_hops = plan.n_hops
    n_pass = max(1, n_complex // (n_hops * (plan.buf_len // 2)))
    hb = synth.power_hops(min(n_pass, 2), n_hops, plan.buf_len, seed=777)
    hb = np.ascontiguousarray(np.tile(hb, (-(-n_pass // hb.shape[0]), 1, 1))[:n_pass])
    win = oracle.port().window_table(wname, 1 << plan.bin_e)
    if kind_pref == "reference" and oracle.have_ref():
        rp = oracle.RefPower()
        with open(os.devnull, "w") as dn:
            saved = os.dup(2)
            os.dup2(dn.fileno(), 2)
            try:
                rp.setup(arg, crop, 1, 0, 0, "rectangle", win)
            finally:
                os.dup2(saved, 2)
                os.close(saved)
        t = rp.time(hb, n_pass, repeats)
        return t, n_pass * n_hops * (plan.buf_len // 2) * repeats, "reference"
    pp = oracle.PowerParams(bin_e=plan.bin_e, buf_len=plan.buf_len)
    t = oracle.port().power_time(pp, win, hb, n_pass, n_hops, repeats)
    return t, n_pass * n_hops * (plan.buf_len // 2) * repeats, "port"


def cpu_leg(workload, cores, target_seconds=12.0):
    """Time the reference C path on `cores` host processes (each its own copy of the globals)."""
    import multiprocessing as mp
    import oracle
    oracle.build()
    kind = "reference" if oracle.have_ref() else "port"
    n_complex = 1 << 23 if workload.startswith("fm") else 1 << 22
    # calibrate on one repeat, then size the run
    t1, n1, kind = _cpu_worker((kind, workload, n_complex, 1))
    repeats = max(1, int(target_seconds / max(t1, 1e-3)))
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    if cores == 1:
        results = [_cpu_worker((kind, workload, n_complex, repeats))]
    else:
        with ctx.Pool(cores) as pool:
            results = pool.map(_cpu_worker, [(kind, workload, n_complex, repeats)] * cores)
    wall = time.perf_counter() - t0
    total = sum(r[1] for r in results)
    slowest = max(r[0] for r in results)
    return {"value": total / slowest / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
            "sample": f"{repeats} x {n1} complex samples of the same workload per core, chunk {CHUNK // 2}; "
                      f"timed inside the C loop (slowest core {slowest:.2f} s, wall {wall:.1f} s)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    leg = cpu_leg(args.workload, cores, target_seconds=max(4.0, 3.0 * (args.steps + args.warmup) / 3))
    res = {"impl": "reference",
           "metric": "I/Q Msamples/s through full_demod() & fix_fft() at 1/2/4/8 B200 vs host CPU",
           "value": leg["value"], "unit": "Msamples/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": (time.perf_counter() - t0) * 1e3 / max(args.steps, 1),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32 C (gcc -O2)",
           "data": "synthetic",
           "config": {"workload": {"fm2b": "rx_fm -M wbfm -s 300k -F 9 -r 48k (2.4 Msps capture -> 48 kHz), reference C path",
                                   }.get(args.workload, args.workload + ", reference C path"),
                      "chunk_complex": CHUNK // 2, "threads": cores},
           "gpu_launches": 0, "cpu_baseline": leg,
           "e2e": {"value": leg["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fm2b", choices=["fm2b", "fm2a", "fm1", "fm5a", "power3", "power4"])
    ap.add_argument("--size-mib", type=int, default=0, help="override the per-GPU input size (testing)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        res = run_reference(args)
        if res is not None:
            print(json.dumps(res))
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (librxb200 has no CPU fallback)")
    rank, local, world = dist_setup(args.gpus)
    if args.workload.startswith("fm"):
        res = run_fm(args, rank, local, world)
    else:
        res = run_power(args, rank, local, world)
    if rank == 0:
        if world == 1 and not args.no_cpu:
            res["cpu_baseline"] = cpu_leg(args.workload, 1, target_seconds=12.0)
        else:
            res["cpu_baseline"] = None
        print(json.dumps(res))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
