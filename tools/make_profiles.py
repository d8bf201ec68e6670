#!/usr/bin/env python
"""Turn one gpurun profiling session (tools/gpu_profile.sh TAG + bench.py > gpurun_out/TAG_bench.json)
into the tracked evidence under profiles/:

    python tools/make_profiles.py r1n            # reads gpurun_out/r1n_*, writes profiles/r1n_* and README.md

  profiles/TAG_ncu_summary.txt   per-kernel `ncu --set full` metrics (duration, DRAM bytes, pipes, stalls)
  profiles/TAG_launches.txt      launch list of bench.py --quick (gpu__time_duration per launch, aggregated)
  profiles/TAG_lines_<chain>.txt top source lines by executed instructions (from --import-source)
  profiles/TAG_bench.json        the bench line of the same build
  profiles/ncu_traffic.json      dram bytes of the dominant kernel (read by bench.py -> roofline.traffic)
  profiles/README.md             the table tying them together
The .ncu-rep files themselves stay in gpurun_out/ (scratch, up to 6 MB each)."""
import collections
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
FS = 2_048_000


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    return dict(zip(names, rows[2])), dict(zip(names, units))


def num(row, units, key, scale=None):
    v = float(row[key].replace(",", ""))
    u = units.get(key, "")
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)
    return v * mult


reps = sorted(glob.glob(os.path.join(G, f"{TAG}_*.ncu-rep")))
with open(os.path.join(P, f"{TAG}_ncu_summary.txt"), "w") as f:
    f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py")] + reps,
                           capture_output=True, text=True).stdout)
for r in reps:
    chain = os.path.basename(r)[len(TAG) + 1:-len(".ncu-rep")]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), r, "25"],
                         capture_output=True, text=True).stdout
    with open(os.path.join(P, f"{TAG}_lines_{chain}.txt"), "w") as f:
        f.write(out)

# launch list -----------------------------------------------------------------------------
launch_txt = ""
lpath = os.path.join(G, f"{TAG}_launches.csv")
if os.path.exists(lpath):
    lines = [l for l in open(lpath) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(t for _, t in agg.values()) or 1.0
    launch_txt = "launch list of `bench.py --steps 4 --warmup 3 --quick` under ncu (gpu__time_duration.sum, ns; cold-cache,\n" \
                 "serialised: compare SHARES with bench.py's call_ms, not absolutes)\n\n"
    for k, (n, t) in agg.items():
        launch_txt += f"{n:4d} launches  {t / n / 1e3:10.1f} us avg  {100 * t / tot:5.1f} % of kernel time   {k}\n"
    with open(os.path.join(P, f"{TAG}_launches.txt"), "w") as f:
        f.write(launch_txt)

# bench line ------------------------------------------------------------------------------
bench = None
bpath = os.path.join(G, f"{TAG}_bench.json")
if os.path.exists(bpath):
    bench = json.load(open(bpath))
    json.dump(bench, open(os.path.join(P, f"{TAG}_bench.json"), "w"), indent=1)

# traffic of the three Rx kernel kinds (bench.py picks the dominant one's entry) ----------------
traffic = {}
for chain, key, streams, sps, what in (
        ("rx_mix", "rx_kernel<AM+SSB, 2048k entry>", 1024, FS, "1024 streams (512 AM + 256 LSB + 256 USB) x 1.000 s"),
        ("rx_fm", "rx_kernel<FM, 2048k entry>", 4096, FS // 2 // 8192 * 8192, "4096 NBFM streams x 0.5 s"),
        ("rx_wbfm", "rx_wbfm_kernel<2048k entry>", 3996, FS // 4 // 8192 * 8192, "3996 WBFM streams x 0.25 s (27 per SM, one untiled launch)")):
    path = os.path.join(G, f"{TAG}_{chain}.ncu-rep")
    if not os.path.exists(path):
        continue
    row, units = raw(path)
    rd, wr = num(row, units, "dram__bytes_read.sum"), num(row, units, "dram__bytes_write.sum")
    traffic[key] = {
        "workload": what + ", 2.048 MS/s entry", "streams": streams, "samples_per_stream": sps,
        "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
        "algorithmic_bytes_per_launch": int(streams * sps * 2.0078125),
        "source": f"profiles/{TAG}_ncu_summary.txt (ncu --set full --clock-control none, gpurun_out/{TAG}_{chain}.ncu-rep): "
                  "dram__bytes_read.sum + dram__bytes_write.sum; bench.py scales it by the unit count of its own launch",
        "note": "above the algorithmic bytes by the halo batches time tiles after the first re-read, the float scratch of "
                "the DC-removal IIR (AM/SSB: 4 B per PCM sample each way) and the state records"}
if traffic:
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)

# README ----------------------------------------------------------------------------------
work = {"rx_mix": ("rx_kernel<AM+SSB,2048k>", 1024 * FS), "rx_iir": ("rx_dc_iir_kernel", 1024 * FS),
        "rx_fm": ("rx_kernel<FM,2048k>", 4096 * (FS // 2 // 8192 * 8192)),
        "rx_wbfm": ("rx_wbfm_kernel<2048k>", 3996 * (FS // 4 // 8192 * 8192)),
        "tx_am": ("tx_kernel<AM>", 4096 * (FS // 4 // 8192 * 8192)), "tx_fm": ("tx_kernel<FM>", 4096 * (FS // 4 // 8192 * 8192)),
        "tx_lsb": ("tx_kernel<SSB>", 4096 * (FS // 4 // 8192 * 8192)), "tx_wbfm": ("tx_wbfm_kernel", 4096 * (FS // 4 // 8192 * 8192))}
peak = 6549.8
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
md = [f"# profiles/ — ncu evidence, round 2 (session `{TAG}`; round 1's files `r1w_*` are kept beside it)", "",
      "B200 (148 SMs, 1965 MHz, no throttle reasons), `ncu --set full --clock-control none --import-source on`, one capture",
      "per chain of the launch after warm-up (`tools/gpu_profile.sh`, `tools/prof_run.py`); summaries made here with",
      "`tools/make_profiles.py` from the `.ncu-rep` files (which stay in `gpurun_out/`). Durations under ncu are",
      "cold-cache and serialised; the bench numbers (CUDA events, un-profiled) are in the last table.", "",
      f"HBM peak used everywhere: **{peak} GB/s, measured** (`MEASURED_PEAKS.json`, copy kernel). Algorithmic bytes: 2.0078125 B per IQ",
      "sample (SURVEY §8d). Issue-slot ceiling of the chip: 148 SM × 4 warp-instr/clk × 1.965 GHz = 1163 G warp-instr/s = 11.4",
      "lane-operations per IQ sample at the HBM roofline.", "",
      "| capture | kernel | samples/launch | duration (ncu) | DRAM read+write | vs algorithmic | DRAM thr. % | issue slots busy % | ALU / FMA / LSU / XU pipe % | warp instr per IQ sample ×32 | top stalls |",
      "|---|---|---|---|---|---|---|---|---|---|---|"]
for r in reps:
    chain = os.path.basename(r)[len(TAG) + 1:-len(".ncu-rep")]
    if chain not in work:
        continue
    row, units = raw(r)
    name, samples = work[chain]
    dur = num(row, units, "gpu__time_duration.sum")
    traffic = num(row, units, "dram__bytes_read.sum") + num(row, units, "dram__bytes_write.sum")
    alg = samples * 2.0078125
    stalls = []
    for k, v in row.items():
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(v.replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    g = lambda k: row.get(k, "nan")
    inst = float(row["smsp__inst_executed.sum"].replace(",", ""))
    md.append(f"| `{TAG}_{chain}` | `{name}` | {samples / 1e9:.3f} G | {dur * 1e3:.3f} ms | {traffic / 1e9:.3f} GB | "
              f"{traffic / alg:.3f}× | {float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | "
              f"{float(g('sm__inst_issued.avg.pct_of_peak_sustained_active')):.1f} | "
              f"{float(g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')):.0f} / "
              f"{float(g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')):.0f} / "
              f"{float(g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active')):.0f} / "
              f"{float(g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')):.0f} | "
              f"{inst * 32 / samples:.1f} | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:3]) + " |")
md += ["", "Reading: every chain issues 8–15 lane-operations per IQ sample against a ceiling of 11.4 at the HBM roofline, so the",
       "kernels are instruction-issue-bound (issue slots 70–90 % busy, `math_pipe_throttle`/`not_selected` the top stalls, DRAM",
       "45–77 %); DRAM traffic is within a few percent of the algorithmic bytes, i.e. nothing is re-read except tile halos.",
       "What changed against round 1 and the experiments behind it: `r2_experiments.md`.", ""]
if launch_txt:
    md += ["## Launch list (`" + f"{TAG}_launches.txt" + "`)", "", "```", launch_txt.rstrip(), "```", ""]
if bench:
    rf = bench["roofline"]
    md += ["## Bench of the same build (`" + f"{TAG}_bench.json" + "`, CUDA events, not under ncu)", "",
           f"* headline (config 5: the mixed-mode batch, 4096 streams × 0.5 s: 1/4 AM, 1/4 NBFM, 1/4 WBFM, 1/8 LSB, 1/8 USB): "
           f"**{bench['value'] / 1e6:.3f} T samples/s**, {bench['ms_per_step']} ms/step = **{rf['step']['frac']:.3f} of the measured HBM peak for the whole step**; "
           f"the kinds run side by side; each kind's own launch (serialised pass): "
           + "; ".join(f"`{k}` {v['streams']} streams {v['launch_ms']} ms = {v['frac']:.3f}" for k, v in rf["kinds"].items()),
           f"* dominant kernel `{rf['kernel']}`: {rf['achieved']} GB/s algorithmic = **{rf['frac']:.3f}**; DRAM traffic {rf.get('traffic')} B per launch "
           f"against {int(rf['algorithmic_bytes_per_launch'])} algorithmic",
           f"* end to end through the C ABI from pinned host memory: {bench['e2e']['value']:.0f} MS/s = {bench['e2e'].get('of_platform_ceiling')} of the box's own "
           f"H2D ceiling ({bench['e2e'].get('platform_h2d_gbs_per_gpu')} GB/s; 2 B per sample in)",
           f"* CPU reference in the same run: {bench.get('cpu_baseline', {}).get('value')} MS/s on {bench.get('cpu_baseline', {}).get('cores')} host threads",
           f"* mixed-mode stream sweep (fraction of the HBM roofline): " + ", ".join(f"{k}: {v['hbm_frac_per_gpu']}" for k, v in bench.get("mixed_mode_stream_sweep", {}).items()),
           f"* WBFM verification: {bench.get('wbfm_tile_fallback_streams')} stream-calls retried, {bench.get('wbfm_serial_rerun_streams')} walked serially in the timed run", "",
           "| chain (4096 streams × 0.5 s) | MS/s | ms | fraction of HBM roofline (of measured) |", "|---|---|---|---|"]
    for k, v in bench.get("modes", {}).items():
        md.append(f"| {k} | {v['MS/s']:.0f} | {v['ms']} | {v['hbm_frac']:.3f} |")
    md.append("")
md += ["## Other files", "",
       "* `r1_ubench_pipes.txt` — instruction-throughput probe (`tools/ubench/pipes.cu`): IMAD/dp2a/PRMT/LOP3/SHF 1.97, add 3.9, FFMA 3.6,",
       "  SHFL 0.99, F2F.F64.F32 0.47, F2I 0.48, I2F 0.99 warp-instr/clk/SM — the numbers behind the pipe-balance and",
       "  \"no double on the hot path\" decisions in DESIGN.md §5.",
       f"* `{TAG}_lines_<chain>.txt` — executed warp instructions per source line (needs `-lineinfo`; inline-PTX lines are unattributed).",
       "* `ncu_traffic.json` — DRAM bytes per launch of the dominant kernel; `bench.py` reports it as `roofline.traffic`.", ""]
open(os.path.join(P, "README.md"), "w").write("\n".join(md))
print("\n".join(md[:40]))
