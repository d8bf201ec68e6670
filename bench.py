#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the HackRfDiags baseband DSP hot path.

    python bench.py --gpus N --steps K --warmup W            (ours: CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's own CPU chain)

Headline workload at every N (weak scaling, per GPU): BASELINE.json configs[4] at its 4k point, the MIXED-MODE
batch -- 4096 streams per GPU, 1/4 AM, 1/4 NBFM, 1/4 WBFM, 1/8 LSB, 1/8 USB, 0.5 s of int8 IQ at 2.048 MS/s
each (4.19 GB in, 16.4 MB of PCM out per step), entered at IqDataProcessor::acceptIqData, streams dealt to
the ranks by hackrfdiags_b200.shard (contiguous ranges of one global plan, no collective).  A step is one
pass over that batch; stream state carries from step to step exactly as consecutive reference calls would.
`value` is input IQ MS/s with the inputs resident in HBM; `e2e` is the same through hrd_rx_process with pinned
HOST buffers (H2D of all IQ and D2H of all PCM inside the timed region).  The same line carries, at every N:
`modes` (each of the eight chains alone, 4096 streams per GPU, with its roofline fraction, the reference's CPU
rate for that chain and a parity record against the reference), `min_mode_hbm_frac`, the 1k..64k mixed-mode
stream sweep sharded over the N GPUs (strong scaling) and the WBFM-modulator stream-count sweep.

Experiment switches (environment; none is set in a driver run): HRD_BENCH_TILE_BATCHES (force the Rx time-tile size),
HRD_BENCH_WBFM_PACK (0 / 1, see wbfm_pack), HRD_BENCH_BACKEND (process-group backend other than nccl),
HRD_BENCH_RANK_MS (every rank prints its own step time to stderr).

oracle/ is used here as the CHECKER only (cpu_baseline / parity legs and --impl reference): the compiled
reference oracle/_ref/libhrd_ref.so (unmodified sources), else the C port oracle/liboracle.so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

FS = 2_048_000
BYTES_PER_IN_SAMPLE_RX = 2.0 + 2.0 / 256.0
BYTES_PER_OUT_SAMPLE_TX = 2.0 + 2.0 / 256.0
MIX = {1: 0.25, 2: 0.25, 3: 0.25, 4: 0.125, 5: 0.125}  # config 5: AM, NBFM, WBFM, LSB, USB
KIND_OF_MODE = {1: 1, 2: 2, 3: 3, 4: 1, 5: 1}          # kernel kind (AM and SSB share a launch)
KIND_KERNEL = {1: "rx_kernel<AM+SSB, 2048k entry>", 2: "rx_kernel<FM, 2048k entry>", 3: "rx_wbfm_kernel<2048k entry>"}

METRIC = ("aggregate input IQ MS/s (config 5: mixed-mode demod batch, 4096 streams/GPU = 1/4 AM + 1/4 NBFM + 1/4 WBFM "
          "+ 1/8 LSB + 1/8 USB, 2.048 MS/s entry)")
UNIT = "MS/s"
MODE_NAMES = {1: "am", 2: "fm", 3: "wbfm", 4: "lsb", 5: "usb"}
EDGE_CLASSES = ("full-range noise", "constant -128", "constant +127", "alternating +-127", "zero")


def ncu_traffic_bytes(kernel, streams, n_samples):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel` on this workload, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written from the .ncu-rep by tools/ncu_summary.py);
    scaled by the unit count when the capture is of the same kernel at another size; None without a capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[kernel]
        return int(t["dram_bytes_per_launch"] * (streams * n_samples) / (t["streams"] * t["samples_per_stream"]))
    except Exception:
        return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------
# synthetic inputs generated on the device (same signal classes as hackrfdiags_b200/synth.py);
# every batch of >= 8 distinct streams ends with one stream of each EDGE class
# ----------------------------------------------------------------------------------------
def make_rx_iq_device(torch, mode, n_distinct, n_samples, device, seed):
    """[n_distinct, 2*n_samples] int8: carrier at -64 kHz, signal of `mode`, amplitude / noise cycled; the last
    five rows are the edge classes (full-range noise, constant -128, constant +127, alternating +-127, zero)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(n_samples, device=device, dtype=torch.float64) / FS
    two_pi = 2.0 * 3.141592653589793
    out = torch.empty((n_distinct, 2 * n_samples), dtype=torch.int8, device=device)
    amps = (20.0, 60.0, 100.0, 127.0)
    sigmas = (1.0, 3.0, 10.0)
    n_edge = 5 if n_distinct >= 8 else 0
    for s in range(n_distinct - n_edge):
        amp, sigma = amps[s % 4], sigmas[(s // 4) % 3]
        if mode == 1:
            env = (1.0 + 0.8 * torch.sin(two_pi * 1000.0 * t)) / 1.8
            ph = torch.zeros_like(t)
        elif mode == 2:
            env = torch.ones_like(t)
            ph = 3.0 * (1.0 - torch.cos(two_pi * 1000.0 * t))
        elif mode == 3:
            env = torch.ones_like(t)
            ph = 45.0 * (1.0 - torch.cos(two_pi * 1000.0 * t)) + 3.0 * (1.0 - torch.cos(two_pi * 10000.0 * t))
        else:
            sign = -1.0 if mode == 4 else 1.0
            a, b = sign * two_pi * 700.0 * t, sign * two_pi * 1900.0 * t
            re, im = 0.5 * (torch.cos(a) + torch.cos(b)), 0.5 * (torch.sin(a) + torch.sin(b))
            env = torch.sqrt(re * re + im * im)
            ph = torch.atan2(im, re)
        ph = ph + two_pi * (-64000.0) * t
        i = amp * env * torch.cos(ph) + sigma * torch.randn(n_samples, device=device, dtype=torch.float64, generator=g)
        q = amp * env * torch.sin(ph) + sigma * torch.randn(n_samples, device=device, dtype=torch.float64, generator=g)
        out[s, 0::2] = torch.clamp(torch.round(i), -128, 127).to(torch.int8)
        out[s, 1::2] = torch.clamp(torch.round(q), -128, 127).to(torch.int8)
    if n_edge:
        e = n_distinct - n_edge
        out[e] = torch.randint(-128, 128, (2 * n_samples,), device=device, generator=g).to(torch.int8)
        out[e + 1] = -128
        out[e + 2] = 127
        alt = torch.full((n_samples,), 127, device=device, dtype=torch.int8)
        alt[1::2] = -127
        out[e + 3, 0::2] = alt
        out[e + 3, 1::2] = alt
        out[e + 4] = 0
    return out


def make_tx_pcm_device(torch, n_distinct, n_samples, device, seed):
    """[n_distinct, n_samples] int16 at 8 kS/s: full-scale sines 300..3400 Hz (reaching -32768), uniform noise,
    amplitude-modulated tone, silence, +-32767 square wave -- cycled by stream index."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.arange(n_samples, device=device, dtype=torch.float64) / 8000.0
    two_pi = 2 * 3.141592653589793
    out = torch.empty((n_distinct, n_samples), dtype=torch.int16, device=device)
    for s in range(n_distinct):
        k = s % 5
        if k == 0:
            x = -32768.0 * torch.cos(two_pi * (300.0 + 97.0 * s % 3100.0) * t)
        elif k == 1:
            x = torch.randint(-32768, 32768, (n_samples,), device=device, generator=g).to(torch.float64)
        elif k == 2:
            x = 12000.0 * torch.sin(two_pi * 440.0 * t) * (0.5 + 0.5 * torch.sin(two_pi * 3.0 * t))
        elif k == 3:
            x = torch.zeros_like(t)
        else:
            x = torch.where(torch.sin(two_pi * (150.0 + 10.0 * s) * t) >= 0, 32767.0, -32767.0).to(torch.float64)
        out[s] = torch.clamp(torch.round(x), -32768, 32767).to(torch.int16)
    return out


def wbfm_pack():
    """HRD_OPT_RX_WBFM_PACK for the bench's batches.  Packing the WBFM launch of a mixed batch onto fewer SMs wins
    2.7 % in a process of its own (1.73 against 1.78 ms per step, the same under torchrun with one rank or with a gloo
    process group) and LOSES 3.3 % once a multi-rank NCCL communicator is alive in the process (1.83 against 1.77 ms,
    both ranks alike; profiles/r2_experiments.md section 5): the kernels of different streams are co-scheduled
    differently then.  So: on, unless this process is one of several NCCL ranks; HRD_BENCH_WBFM_PACK overrides."""
    if os.environ.get("HRD_BENCH_WBFM_PACK"):
        return int(os.environ["HRD_BENCH_WBFM_PACK"])
    nccl_ranks = int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("HRD_BENCH_BACKEND", "nccl") == "nccl"
    return 0 if nccl_ranks else 1


def tile_rows(torch, distinct, n_rows):
    reps = (n_rows + distinct.shape[0] - 1) // distinct.shape[0]
    return distinct.repeat(reps, 1)[:n_rows].contiguous()


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------
def time_calls(torch, calls, steps, warmup):
    """calls: list of zero-arg launchers.  Returns (ms_per_step, [ms per call])."""
    for _ in range(warmup):
        for c in calls:
            c()
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(calls) + 1)] for _ in range(steps)]
    for k in range(steps):
        ev[k][0].record()
        for j, c in enumerate(calls):
            c()
            ev[k][j + 1].record()
    torch.cuda.synchronize()
    total = ev[0][0].elapsed_time(ev[-1][-1])
    per_call = [sum(ev[k][j].elapsed_time(ev[k][j + 1]) for k in range(steps)) / steps for j in range(len(calls))]
    return total / steps, per_call


def make_rx_batch(torch, capi, device, groups, n_samples, seed):
    """ONE batch holding every group's streams: groups = [(mode, n_streams)]; returns (batch, iq, pcm, distinct)
    with distinct[mode] = the rows the group's streams repeat."""
    n = sum(g[1] for g in groups)
    iq = torch.empty((n, 2 * n_samples), dtype=torch.int8, device=device)
    b = capi.Batch(n, capi.RX, device.index or 0)
    if os.environ.get("HRD_BENCH_TILE_BATCHES"):  # experiments (tools/prof_run.py): force the time-tile size
        b.set_option(capi.OPT_RX_TILE_BATCHES, int(os.environ["HRD_BENCH_TILE_BATCHES"]))
    b.set_option(capi.OPT_RX_WBFM_PACK, wbfm_pack())
    at = 0
    n_distinct = {}
    for mode, cnt in groups:
        distinct = make_rx_iq_device(torch, mode, min(cnt, 32), n_samples, device, seed + mode)
        iq[at:at + cnt] = tile_rows(torch, distinct, cnt)
        n_distinct[mode] = (at, distinct.shape[0], cnt)
        del distinct
        for s in range(at, at + cnt):
            b.set_mode(mode, s)
        at += cnt
    pcm = torch.zeros((n, n_samples // 256), dtype=torch.int16, device=device)
    return b, iq, pcm, n_distinct


def repeats_identical(torch, out, layout):
    """Rows that got identical input must give identical output: EVERY repeat of every distinct row, on the device."""
    bad = 0
    for at, nd, cnt in layout.values():
        full = cnt // nd
        if full >= 2:
            v = out[at:at + full * nd].view(full, nd, -1)
            bad += int((v != v[0:1]).any(dim=2).sum().item())
        rest = cnt - full * nd
        if rest and full:
            bad += int((out[at + full * nd:at + cnt] != out[at:at + rest]).any(dim=1).sum().item())
    return bad


def bench_rx_modes(torch, capi, device, groups, n_samples, steps, warmup, seed, serial_pass=False):
    """One batch, one hrd_rx_process call per step.  Returns a dict: ms (per step), kernel_ms (the tile kernels'
    span), tail_ms (the AM/SSB IIR pass), both from CUDA events the library records on the launching stream
    inside the timed steps (HRD_OPT_PROFILE); launches per step; repeat_mismatches (all row repeats compared on the
    device); and, with serial_pass, kind_ms: each kernel kind's own duration from three extra steps with the
    kinds run one after the other (HRD_OPT_RX_SERIAL) -- side by side their spans overlap and cannot be told apart."""
    stream = torch.cuda.current_stream().cuda_stream
    b, iq, pcm, layout = make_rx_batch(torch, capi, device, groups, n_samples, seed)
    b.set_option(capi.OPT_PROFILE, 1)
    call = lambda: b.rx_device(iq.data_ptr(), iq.shape[1], iq.stride(0), pcm.data_ptr(), pcm.stride(0),
                               capi.ENTRY_2048K, stream)
    l0 = b.launch_count()
    ms_step, _ = time_calls(torch, [call], steps, warmup)
    launches = (b.launch_count() - l0) // (steps + warmup)
    k = min(steps, 32)
    res = {"ms": ms_step, "kernel_ms": sum(b.kernel_ms(0, a) for a in range(k)) / k,
           "tail_ms": sum(b.kernel_ms(1, a) for a in range(k)) / k, "launches": launches,
           "repeat_mismatches": repeats_identical(torch, pcm, layout),
           "wbfm_fallbacks": b.wbfm_fallback_count(), "wbfm_serial": b.wbfm_serial_count()}
    if serial_pass:
        b.set_option(capi.OPT_RX_SERIAL, 1)
        for _ in range(4):
            call()
        torch.cuda.synchronize()
        kinds = sorted({KIND_OF_MODE[m] for m, _ in groups})
        res["kind_ms"] = {kd: sum(b.kernel_ms(10 + kd, a) for a in range(3)) / 3 for kd in kinds}
        b.set_option(capi.OPT_RX_SERIAL, 0)
    return res, (b, iq, pcm, layout)


def bench_tx_mode(torch, capi, device, mode, n, n_pcm, steps, warmup, seed, want=False):
    stream = torch.cuda.current_stream().cuda_stream
    distinct = make_tx_pcm_device(torch, min(n, 32), n_pcm, device, seed)
    pcm = tile_rows(torch, distinct, n)
    iq = torch.empty((n, n_pcm * 512), dtype=torch.int8, device=device)
    b = capi.Batch(n, capi.TX, device.index or 0)
    b.set_mode(mode)
    call = lambda: b.tx_device(pcm.data_ptr(), n_pcm, pcm.stride(0), iq.data_ptr(), iq.stride(0), stream)
    ms_step, per_call = time_calls(torch, [call], steps, warmup)
    if want:
        bad = repeats_identical(torch, iq, {mode: (0, distinct.shape[0], n)})
        return ms_step, bad, distinct
    return ms_step


def reduce_max_ms(torch, dist, device, values):
    """max over ranks of a list of per-rank times (identity on one rank)"""
    if not dist:
        return list(values)
    t = torch.tensor(list(values), device=device if dist.get_backend() == "nccl" else "cpu", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def run_ours(args):
    import torch
    from hackrfdiags_b200 import capi, shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        # NCCL prints a version banner on STDOUT when its debug level is VERSION (environment or nccl.conf), in
        # front of the one JSON line the contract asks for.  Ask for warnings only unless the caller chose a
        # level, and send whatever the library still writes to fd 1 while the communicator comes up to stderr.
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            if os.environ.get("HRD_BENCH_BACKEND", "nccl") == "nccl":
                dist.init_process_group("nccl", device_id=device)
            else:  # experiments only: the control plane over another backend (the reductions then go through the host)
                dist.init_process_group(os.environ["HRD_BENCH_BACKEND"])
            dist.barrier()  # the communicator (and its banner) is created here at the latest
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    peak, peak_src = measured_peak_gbs()

    # the rank's share of ONE global plan of world * streams mixed-mode streams (weak scaling: the plan grows with N)
    n_samples = int(args.seconds * FS) // 8192 * 8192
    plan = shard.mixed_mode_plan(world * args.streams, MIX)
    mine = shard.shard_modes(plan, world, rank)
    groups = shard.mode_groups(mine)
    my_streams = len(mine)
    in_samples_per_step = my_streams * n_samples

    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    r, keep = bench_rx_modes(torch, capi, device, groups, n_samples, args.steps, args.warmup, seed=1234 + rank,
                             serial_pass=True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    if os.environ.get("HRD_BENCH_RANK_MS"):  # experiments: every rank's own step time
        print(f"rank {rank}: {r['ms']:.4f} ms per step, tile kernels {r['kernel_ms']:.4f}", file=sys.stderr, flush=True)
    ms_max = reduce_max_ms(torch, dist, device, [r["ms"]])[0]
    clocks = sampler.stop() if sampler else None
    value = world * in_samples_per_step / (ms_max * 1e-3) / 1e6

    # roofline of the dominant kernel: the kind that takes longest when the kinds run one after the other
    per_kind_streams = {}
    for m, c in groups:
        per_kind_streams[KIND_OF_MODE[m]] = per_kind_streams.get(KIND_OF_MODE[m], 0) + c
    dom = max(r["kind_ms"], key=lambda kd: r["kind_ms"][kd])
    dom_bytes = per_kind_streams[dom] * n_samples * BYTES_PER_IN_SAMPLE_RX
    achieved = dom_bytes / (r["kind_ms"][dom] * 1e-3) / 1e9
    step_bytes = in_samples_per_step * BYTES_PER_IN_SAMPLE_RX
    roofline = {"bound": "hbm", "kernel": KIND_KERNEL[dom], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4),
                "traffic": ncu_traffic_bytes(KIND_KERNEL[dom], per_kind_streams[dom], n_samples), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": round(r["kind_ms"][dom], 4),
                "how": "the three kernel kinds of the mixed batch run side by side in the timed steps; each kind's own "
                       "launch time comes from three extra steps with the kinds serialised (HRD_OPT_RX_SERIAL), CUDA "
                       "events on the launching stream",
                "kinds": {KIND_KERNEL[kd]: {"streams": per_kind_streams[kd], "launch_ms": round(ms, 4),
                                            "frac": round(per_kind_streams[kd] * n_samples * BYTES_PER_IN_SAMPLE_RX
                                                          / (ms * 1e-3) / 1e9 / peak, 4)}
                          for kd, ms in r["kind_ms"].items()},
                "step": {"algorithmic_bytes": step_bytes, "ms": round(r["ms"], 4),
                         "frac": round(step_bytes / (r["ms"] * 1e-3) / 1e9 / peak, 4)}}

    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_max, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 (Q15 accumulate over int8/int16 samples) + f32 detector / IIR sections",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[4] (mixed-mode) at 4096 streams/GPU: 1/4 AM, 1/4 NBFM, 1/4 WBFM, 1/8 LSB, "
                               f"1/8 USB, {n_samples / FS:.3f} s of int8 IQ @2.048 MS/s each, IqDataProcessor entry; every "
                               "group of 32 distinct streams ends with the five edge classes",
                   "streams_per_gpu": args.streams, "input_bytes_per_step_per_gpu": 2 * in_samples_per_step,
                   "l2": "inputs per step exceed the 126 MB L2 many times over; no flush needed",
                   "sharding": "one global plan, contiguous stream ranges per rank (hackrfdiags_b200.shard), no collective",
                   "wbfm_pack": wbfm_pack(),
                   "mode_groups_rank0": [[MODE_NAMES[m], c] for m, c in groups]},
        "roofline": roofline, "gpu_launches": r["launches"] * args.steps * world,
        "call_ms": {"tile_kernels": round(r["kernel_ms"], 4), "iir_tail": round(r["tail_ms"], 4)},
        "repeat_mismatches": r["repeat_mismatches"], "wbfm_tile_fallback_streams": r["wbfm_fallbacks"],
        "wbfm_serial_rerun_streams": r["wbfm_serial"],
    }
    if clocks:
        out["clocks"] = clocks
    del keep
    torch.cuda.empty_cache()

    # end to end through the C ABI with host buffers: every rank drives its own GPU at the same time;
    # whole-job value = all ranks' samples / slowest rank's time
    e2e = run_e2e(torch, capi, device, args, groups, n_samples, dist)
    if rank == 0:
        out["e2e"] = e2e
    if not args.quick:
        modes = run_mode_sweep(torch, capi, device, args, peak, dist, rank)
        sweep = run_stream_sweep(torch, capi, shard, device, args, peak, dist, rank)
        txsweep = run_tx_wbfm_sweep(torch, capi, device, args, peak) if rank == 0 else None
        adapters = run_adapters(torch, capi, device, args) if rank == 0 else None
        if rank == 0:
            out["modes"] = modes
            chains = [k for k in modes if k.split("_")[0] in ("rx", "tx") and k.count("_") == 1]
            out["min_mode_hbm_frac"] = min((modes[k]["hbm_frac"], k) for k in chains)
            out["mixed_mode_stream_sweep"] = sweep
            out["tx_wbfm_stream_sweep"] = txsweep
            out["adapters"] = adapters
            out["parity"] = {k: v["parity"] for k, v in modes.items() if "parity" in v}
            out["cpu_baseline"] = cpu_baseline(args)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def run_e2e(torch, capi, device, args, groups, n_samples, dist):
    """Same workload through the C ABI with pinned HOST buffers: H2D + kernels + D2H per step, on every rank."""
    world = dist.get_world_size() if dist else 1
    steps = max(2, min(args.steps, 5))
    b, iq, pcm, _ = make_rx_batch(torch, capi, device, groups, n_samples, 99)
    host_iq = torch.empty(iq.shape, dtype=torch.int8, pin_memory=True)
    host_iq.copy_(iq)
    host_pcm = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
    n_streams = iq.shape[0]
    del iq, pcm
    torch.cuda.synchronize()

    def step():  # returns only when the PCM is in host memory
        b.rx_host_ptr(host_iq.data_ptr(), host_iq.shape[1], host_iq.stride(0), host_pcm.data_ptr(), host_pcm.stride(0))

    step()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    dt = reduce_max_ms(torch, dist, device, [dt])[0]
    value = world * n_streams * n_samples / dt / 1e6
    # the platform's ceiling for this step: the same pinned buffer copied to the device by a plain cudaMemcpyAsync
    # and nothing else, every rank at the same time (the host side -- root complexes, memory channels -- is shared)
    del b
    dev_iq = torch.empty(host_iq.shape, dtype=torch.int8, device=device)
    dev_iq.copy_(host_iq, non_blocking=True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    dt_copy = None
    for _ in range(2):  # the better of two passes: it is a ceiling
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            dev_iq.copy_(host_iq, non_blocking=True)
        torch.cuda.synchronize()
        dt_pass = reduce_max_ms(torch, dist, device, [(time.perf_counter() - t0) / steps])[0]
        dt_copy = dt_pass if dt_copy is None else min(dt_copy, dt_pass)
    del dev_iq
    return {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": host_iq.numel() * world,
            "d2h_bytes_per_step": host_pcm.numel() * 2 * world, "ms_per_step": round(dt * 1e3, 3), "steps": steps,
            "h2d_gbs_per_gpu": round(host_iq.numel() / dt / 1e9, 2),
            "platform_h2d_gbs_per_gpu": round(host_iq.numel() / dt_copy / 1e9, 2),
            "of_platform_ceiling": round(dt_copy / dt, 4),
            "note": "hrd_rx_process(HRD_MEM_HOST) on pinned host buffers, one call per step per GPU, copies inside "
                    "the call; whole job = all ranks, slowest rank's wall time; bound by PCIe H2D (2 B per IQ sample). "
                    "platform_h2d_gbs_per_gpu = the same buffers through a bare cudaMemcpyAsync on all ranks at once: "
                    "what this box's host side delivers per GPU at this N"}


# ----------------------------------------------------------------------------------------
# checkers (test infrastructure): the compiled reference when it travelled, else the C port
# ----------------------------------------------------------------------------------------
_checker = None


def checker():
    global _checker
    if _checker is None:
        import cpu_checkers
        if cpu_checkers.have_ref():
            _checker = ("reference", cpu_checkers.Ref())
        else:
            _checker = ("port", cpu_checkers.Oracle())
    return _checker


def cpu_rx(mode, iq, cores):
    """(seconds, pcm[n, n_pcm]) of the checker's Rx chain over iq[n, bytes] on `cores` threads (1 for the port)."""
    import numpy as np
    kind, c = checker()
    if kind == "reference":
        return c.bench_rx(mode, iq, cores, want_pcm=True)
    t0 = time.perf_counter()
    pcm = np.stack([c.run_rx(mode, row) for row in iq])
    return time.perf_counter() - t0, pcm


def cpu_tx(mode, pcm, cores):
    import numpy as np
    kind, c = checker()
    if kind == "reference":
        return c.bench_tx(mode, pcm, cores, want_iq=True)
    t0 = time.perf_counter()
    iq = np.stack([c.run_tx(mode, row) for row in pcm])
    return time.perf_counter() - t0, iq


def parity_record(got, want, tol, note=""):
    import numpy as np
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    return {"checker": checker()[0], "streams": int(got.shape[0]), "samples_per_stream": int(got.shape[1]),
            "max_abs_err": int(d.max()) if d.size else 0, "mismatches": int((d > 0).sum()),
            "tolerance": tol, "ok": bool((d.max() if d.size else 0) <= tol), "note": note}


def rx_parity_and_cpu(torch, capi, device, mode, distinct_rows, cores):
    """The distinct rows of a mode's batch (edge classes included), first 0.25 s: CUDA path from a fresh batch
    through the C ABI (host buffers) against the checker; the checker's run is also the chain's CPU rate."""
    import numpy as np
    host = distinct_rows.cpu().numpy()
    b = capi.Batch(host.shape[0], capi.RX, device.index or 0)
    b.set_mode(mode)
    got = b.rx(host)
    b.close()
    reps = max(1, (2 * cores + host.shape[0] - 1) // host.shape[0])
    cpu_in = np.ascontiguousarray(np.tile(host, (reps, 1)))
    cpu_rx(mode, cpu_in[:host.shape[0]], cores)  # warm
    dt, want = cpu_rx(mode, cpu_in, cores)
    cpu = {"MS/s": round(cpu_in.shape[0] * (cpu_in.shape[1] // 2) / dt / 1e6, 1), "cores": cores if checker()[0] == "reference" else 1,
           "kind": checker()[0], "sample": f"{cpu_in.shape[0]} streams x {cpu_in.shape[1] / 2 / FS:.3f} s"}
    return parity_record(got, want[:host.shape[0]], 0, "bit-exact required; rows end with: " + ", ".join(EDGE_CLASSES)), cpu


def tx_parity_and_cpu(torch, capi, device, mode, distinct_rows, cores):
    import numpy as np
    host = distinct_rows.cpu().numpy()
    b = capi.Batch(host.shape[0], capi.TX, device.index or 0)
    b.set_mode(mode)
    got = b.tx(host)
    b.close()
    reps = max(1, (2 * cores + host.shape[0] - 1) // host.shape[0])
    cpu_in = np.ascontiguousarray(np.tile(host, (reps, 1)))
    cpu_tx(mode, cpu_in[:host.shape[0]], cores)
    dt, want = cpu_tx(mode, cpu_in, cores)
    cpu = {"MS/s": round(cpu_in.shape[0] * cpu_in.shape[1] * 256 / dt / 1e6, 1), "cores": cores if checker()[0] == "reference" else 1,
           "kind": checker()[0], "sample": f"{cpu_in.shape[0]} streams x {cpu_in.shape[1] / 8000:.3f} s"}
    tol = 0  # FM included: libm's cosf / sinf are restated bit for bit
    return parity_record(got, want[:host.shape[0]], tol, "int8 I,Q; sines reaching -32768, noise, AM tone, silence, square wave"), cpu


def run_mode_sweep(torch, capi, device, args, peak, dist, rank):
    """Every chain alone, sweep_streams streams per GPU on every rank: MS/s (whole job, slowest rank's time), fraction
    of the HBM roofline, and on rank 0 the chain's CPU rate and a parity record against the checker."""
    world = dist.get_world_size() if dist else 1
    cores = os.cpu_count() or 1
    res = {}
    n_streams = args.sweep_streams
    n_samples = int(args.sweep_seconds * FS) // 8192 * 8192
    n_par = min(n_samples, int(0.25 * FS) // 131072 * 131072)
    for mode in (1, 2, 3, 4, 5):
        r, keep = bench_rx_modes(torch, capi, device, [(mode, n_streams)], n_samples, 5, 3, seed=7 + rank)
        ms, kms, tms = reduce_max_ms(torch, dist, device, [r["ms"], r["kernel_ms"], r["tail_ms"]])
        sps = world * n_streams * n_samples / (ms * 1e-3)
        row = {"streams_per_gpu": n_streams, "MS/s": round(sps / 1e6, 1), "ms": round(ms, 3),
               "hbm_frac": round(sps / world * BYTES_PER_IN_SAMPLE_RX / 1e9 / peak, 4),
               "kernel_ms": round(kms, 3), "tail_ms": round(tms, 3), "repeat_mismatches": r["repeat_mismatches"]}
        if rank == 0:
            _, iq, _, layout = keep
            at, nd, _ = layout[mode]
            try:
                row["parity"], row["cpu"] = rx_parity_and_cpu(torch, capi, device, mode, iq[at:at + nd, :2 * n_par], cores)
            except Exception as e:  # the checker libraries are missing: say so, substitute nothing
                row["parity"] = {"ok": None, "note": f"checker unavailable: {e}"}
        res[f"rx_{MODE_NAMES[mode]}"] = row
        del keep
        torch.cuda.empty_cache()
    n_pcm = n_samples // 256
    for mode in (1, 2, 3, 4, 5):
        ms, bad, distinct = bench_tx_mode(torch, capi, device, mode, n_streams, n_pcm, 5, 3, seed=11 + rank, want=True)
        ms = reduce_max_ms(torch, dist, device, [ms])[0]
        sps = world * n_streams * n_pcm * 256 / (ms * 1e-3)
        row = {"streams_per_gpu": n_streams, "MS/s": round(sps / 1e6, 1), "ms": round(ms, 3),
               "hbm_frac": round(sps / world * BYTES_PER_OUT_SAMPLE_TX / 1e9 / peak, 4), "repeat_mismatches": bad}
        if rank == 0:
            try:
                row["parity"], row["cpu"] = tx_parity_and_cpu(torch, capi, device, mode, distinct[:, :n_par // 256], cores)
            except Exception as e:
                row["parity"] = {"ok": None, "note": f"checker unavailable: {e}"}
        res[f"tx_{MODE_NAMES[mode]}"] = row
        del distinct
        torch.cuda.empty_cache()
    if rank == 0:
        res.update(run_next_rows(torch, capi, device, args, peak))
    return res


def run_next_rows(torch, capi, device, args, peak):
    """SURVEY 8f rows, measured like the chains above (sweep_streams x sweep_seconds, device-resident, rank 0):
    the squelched receive call (AM streams whose level closes the gate on every other pair of 64 ms blocks) and the
    signals/ tool chain on the transmit side."""
    res = {}
    n_streams = args.sweep_streams
    n_samples = int(args.sweep_seconds * FS) // 131072 * 131072
    stream = torch.cuda.current_stream().cuda_stream
    # squelch: bursty level (loud, loud, quiet, quiet, ...) so that the tracker opens, holds its tail and closes
    b, iq, pcm, _ = make_rx_batch(torch, capi, device, [(1, n_streams)], n_samples, seed=17)
    blocks = n_samples // 131072
    for k in range(blocks):
        if (k // 2) % 2 == 1:
            iq[:, k * 262144:(k + 1) * 262144].div_(32, rounding_mode="floor")
    b.set_param(capi.PARAM_SQUELCH_THRESHOLD, -40.0)
    call = lambda: b.rx_device(iq.data_ptr(), iq.shape[1], iq.stride(0), pcm.data_ptr(), pcm.stride(0), capi.ENTRY_2048K, stream)
    ms, _ = time_calls(torch, [call], 5, 3)
    _, allowed = b.squelch_report()
    sps = n_streams * n_samples / (ms * 1e-3)
    res["rx_am_squelched"] = {"streams": n_streams, "MS/s": round(sps / 1e6, 1), "ms": round(ms, 3),
                              "hbm_frac": round(sps * BYTES_PER_IN_SAMPLE_RX / 1e9 / peak, 4),
                              "blocks_per_call": int(blocks), "open_fraction": round(float(allowed.mean()), 3)}
    del b, iq, pcm
    torch.cuda.empty_cache()
    n_pcm = n_samples // 256
    for name, mode in (("tx_sig_pm", capi.MODE_PM), ("tx_sig_dsb", capi.MODE_DSB)):
        ms = bench_tx_mode(torch, capi, device, mode, n_streams, n_pcm, 5, 3, seed=19)
        torch.cuda.empty_cache()
        sps = n_streams * n_pcm * 256 / (ms * 1e-3)
        res[name] = {"streams": n_streams, "MS/s": round(sps / 1e6, 1), "ms": round(ms, 3),
                     "hbm_frac": round(sps * BYTES_PER_OUT_SAMPLE_TX / 1e9 / peak, 4)}
    return res


def run_stream_sweep(torch, capi, shard, device, args, peak, dist, rank):
    """BASELINE configs[4]: mixed-mode jobs of 1k .. 64k streams IN TOTAL, dealt to the N ranks by shard.shard_modes
    (strong scaling: the job is fixed, each GPU owns 1/N of the streams), the same total signal per point."""
    world = dist.get_world_size() if dist else 1
    res = {}
    total_samples = 4096 * (int(0.5 * FS) // 8192 * 8192)
    for n_total in (1024, 4096, 16384, 65536):
        n_samples = max(8192, total_samples // n_total // 8192 * 8192)
        plan = shard.mixed_mode_plan(n_total, MIX)
        mine = shard.shard_modes(plan, world, rank)
        groups = shard.mode_groups(mine)
        r, keep = bench_rx_modes(torch, capi, device, groups, n_samples, 5, 3, seed=13 + rank)
        del keep
        torch.cuda.empty_cache()
        ms = reduce_max_ms(torch, dist, device, [r["ms"]])[0]
        sps = n_total * n_samples / (ms * 1e-3)
        res[str(n_total)] = {"streams_total": n_total, "streams_rank0": len(mine), "seconds_per_stream": round(n_samples / FS, 4),
                             "MS/s": round(sps / 1e6, 1), "ms": round(ms, 3),
                             "hbm_frac_per_gpu": round(sps / world * BYTES_PER_IN_SAMPLE_RX / 1e9 / peak, 4),
                             "launches_per_step": r["launches"], "repeat_mismatches": r["repeat_mismatches"]}
    return res


def run_adapters(torch, capi, device, args):
    """SURVEY 8f row 2 at rate: producer threads push 262144-byte blocks of every stream into the page-locked pool,
    the consumer keeps three rounds in flight (hrd_rx_pipe_*).  A round is one 64 ms block of every stream, so real
    time needs 15.625 rounds/s; reported: rounds/s sustained, the H2D rate that is, and how it compares."""
    import ctypes as C
    import threading
    lib = capi.load()
    vp = C.c_void_p
    lib.hrd_rx_pipe_create.argtypes = [vp, vp, C.c_int, C.POINTER(vp)]
    lib.hrd_rx_pipe_destroy.argtypes = [vp]
    lib.hrd_rx_pipe_submit.argtypes = [vp]
    lib.hrd_rx_pipe_collect.argtypes = [vp, vp, vp, vp]
    lib.hrd_iq_queue_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hrd_iq_queue_push_rows.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, vp, C.c_size_t, C.c_uint32]
    lib.hrd_iq_queue_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32)]
    lib.hrd_iq_queue_destroy.argtypes = [vp]
    from hackrfdiags_b200 import shard
    res = {}
    import numpy as np
    rounds, n_prod = 24, 6
    for n in (1024, 2048):
        plan = shard.mixed_mode_plan(n, MIX)
        b = capi.Batch(n, capi.RX, device.index or 0)
        for s, m in enumerate(plan):
            b.set_mode(m, s)
        src = make_rx_iq_device(torch, 1, 32, 131072, device, 31).cpu().numpy()  # one block per distinct row
        q, pipe = vp(), vp()
        if lib.hrd_iq_queue_create(n, C.byref(q)) or lib.hrd_rx_pipe_create(b.h, q, 3, C.byref(pipe)):
            res[str(n)] = {"error": "allocation failed"}
            continue

        rows_of = {}
        for i in range(n_prod):  # each producer's block of every one of its streams, as one row matrix
            lo, hi = i * n // n_prod, (i + 1) * n // n_prod
            rows_of[lo] = np.ascontiguousarray(np.stack([src[s % 32] for s in range(lo, hi)]))

        def producer(lo, hi):
            st = (C.c_uint32 * 3)()
            mine = rows_of[lo]
            for k in range(rounds):
                while True:
                    lib.hrd_iq_queue_stats(q, hi - 1, st)
                    if st[0] < 8:
                        break
                lib.hrd_iq_queue_push_rows(q, lo, hi - lo, k, mine.ctypes.data, mine.strides[0], 262144)

        threads = [threading.Thread(target=producer, args=(i * n // n_prod, (i + 1) * n // n_prod)) for i in range(n_prod)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        done = 0
        while done < rounds:
            while lib.hrd_rx_pipe_submit(pipe) == 1:
                pass
            if lib.hrd_rx_pipe_collect(pipe, None, None, None) == 1:
                done += 1
        dt = time.perf_counter() - t0
        for t in threads:
            t.join()
        lib.hrd_rx_pipe_destroy(pipe)
        lib.hrd_iq_queue_destroy(q)
        del b
        rps = rounds / dt
        res[str(n)] = {"rounds_per_s": round(rps, 1), "real_time_need": 15.625, "times_real_time": round(rps / 15.625, 2),
                       "h2d_gbs": round(n * 262144 * rps / 1e9, 2), "MS/s": round(n * 131072 * rps / 1e6, 1),
                       "producer_threads": n_prod, "pipe_depth": 3, "rounds": rounds}
    res["note"] = ("hrd_iq_queue_push_rows by producer threads (each block pushed is the reference's memcpy into the pool) + hrd_rx_pipe_submit / "
                   "collect on the consumer; the producers' memcpy into the pool shares the wall time")
    return res


def run_tx_wbfm_sweep(torch, capi, device, args, peak):
    """The WBFM modulator over the stream count (one GPU): its NCO phase chain is serial per stream (256 000 dependent
    steps per stream-second), so below a few thousand streams that chain, not the bandwidth, sets the time."""
    res = {}
    n_pcm = (int(0.5 * FS) // 8192 * 8192) // 256
    for n in (256, 1024, 4096, 16384):
        ms = bench_tx_mode(torch, capi, device, 3, n, n_pcm, 5, 3, seed=23)
        torch.cuda.empty_cache()
        sps = n * n_pcm * 256 / (ms * 1e-3)
        res[str(n)] = {"MS/s": round(sps / 1e6, 1), "ms": round(ms, 3),
                       "hbm_frac": round(sps * BYTES_PER_OUT_SAMPLE_TX / 1e9 / peak, 4)}
    return res


# ----------------------------------------------------------------------------------------
# CPU arms (the compiled reference, oracle/_ref/libhrd_ref.so; else the C port)
# ----------------------------------------------------------------------------------------
def cpu_sample(n_streams_total, seconds):
    """A bounded sample of the headline workload: the same mode mix, fewer streams."""
    import numpy as np
    from hackrfdiags_b200 import shard, synth
    n_samples = int(seconds * FS) // 8192 * 8192
    plan = shard.mixed_mode_plan(n_streams_total, MIX)
    parts = []
    for mode in sorted(set(plan)):
        k = plan.count(mode)
        distinct = np.stack([synth.rx_stream(mode, n_samples, stream=s) for s in range(min(k, 4))])
        parts.append((mode, np.ascontiguousarray(np.tile(distinct, ((k + 3) // 4, 1))[:k])))
    return parts, n_samples


def time_cpu(parts, n_samples, cores):
    t = 0.0
    for mode, iq in parts:
        dt, _ = cpu_rx(mode, iq, cores)
        t += dt
    n = sum(iq.shape[0] for _, iq in parts)
    return n * n_samples / t / 1e6, t


def cpu_sample_size(cores):
    return (16 * cores, 0.5) if checker()[0] == "reference" else (10, 0.125)


def cpu_baseline(args):
    cores = os.cpu_count() or 1
    try:
        kind = checker()[0]
        n, secs = cpu_sample_size(cores)
        parts, n_samples = cpu_sample(n, secs)
        time_cpu(parts, n_samples, cores)  # warm
        value, t = time_cpu(parts, n_samples, cores)
        used = cores if kind == "reference" else 1
        return {"value": round(value, 1), "unit": UNIT, "cores": used, "kind": kind,
                "sample": f"{sum(p[1].shape[0] for p in parts)} streams x {n_samples / FS:.3f} s, the headline's mode mix, "
                          + ("unmodified reference classes (oracle/_ref), one object graph per stream, " if kind == "reference"
                             else "the C port (oracle/liboracle.so), ")
                          + f"{used} threads, {t:.2f} s wall"}
    except Exception as e:  # the checker library is missing: report, do not substitute anything
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": f"unavailable: {e}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = checker()[0]
    used = cores if kind == "reference" else 1
    n, secs = cpu_sample_size(cores)
    parts, n_samples = cpu_sample(n, secs)
    for _ in range(max(1, min(args.warmup, 2))):
        time_cpu(parts, n_samples, cores)
    t_tot, n_tot = 0.0, 0
    for _ in range(args.steps):
        v, t = time_cpu(parts, n_samples, cores)
        t_tot += t
        n_tot += sum(p[1].shape[0] for p in parts) * n_samples
    value = n_tot / t_tot / 1e6
    sample = (f"each step: {sum(p[1].shape[0] for p in parts)} streams x {n_samples / FS:.3f} s of the same "
              f"AM/NBFM/WBFM/LSB/USB mix through "
              + ("the unmodified reference classes (oracle/_ref/libhrd_ref.so), " if kind == "reference" else "the C port (oracle/liboracle.so), ")
              + f"{used} host threads")
    out = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT,
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(t_tot / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "int16 Q15 / f32 (reference CPU arithmetic)", "data": "synthetic",
           # the product arm's workload, named the same way; each step times a bounded sample of it
           "config": {"workload": "BASELINE configs[4] (mixed-mode) at 4096 streams/GPU: 1/4 AM, 1/4 NBFM, 1/4 WBFM, 1/8 LSB, "
                                  f"1/8 USB, int8 IQ @2.048 MS/s, IqDataProcessor entry",
                      "streams_per_gpu": args.streams, "bounded_sample": "see cpu_baseline.sample"},
           "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": used, "kind": kind,
                            "sample": sample},
           "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=4096, help="streams per GPU (headline: the mixed-mode batch)")
    ap.add_argument("--seconds", type=float, default=0.5, help="signal seconds per stream per step")
    ap.add_argument("--sweep-streams", type=int, default=4096)
    ap.add_argument("--sweep-seconds", type=float, default=0.5)
    ap.add_argument("--quick", action="store_true", help="skip the mode table, the sweeps, parity and the CPU baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
