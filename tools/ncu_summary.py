#!/usr/bin/env python
"""Summarise .ncu-rep captures into the small text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/r1a_rx_mix.ncu-rep [more.ncu-rep ...] > profiles/r1_ncu_summary.txt

Reads each report with `ncu -i ... --page raw --csv` (works without a GPU) and prints the
metrics the roofline argument needs: duration, DRAM bytes, DRAM/L2/L1 throughput %, issue-slot
utilisation, the pipe mix, occupancy, registers, and the top warp-stall reasons."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % (dram__)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy % (inst_issued)"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp issue active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe fma %"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe fmaheavy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe fp64 %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe lsu %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe xu %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu cycles active %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma cycles active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__cycles_active.avg", "smsp cycles active"),
    ("sm__cycles_elapsed.max", "sm cycles elapsed"),
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    return [dict(zip(names, r)) for r in rows[2:]], dict(zip(names, units))


def main():
    for path in sys.argv[1:]:
        launches, units = load(path)
        for row in launches:
            print(f"=== {path}")
            print(f"kernel: {row.get('Kernel Name')}  grid {row.get('Grid Size')} block {row.get('Block Size')}")
            for key, label in KEYS:
                if key in row and row[key] != "":
                    print(f"  {label:34s} {row[key]:>18s} {units.get(key, '')}   [{key}]")
            try:
                rd = float(row["dram__bytes_read.sum"].replace(",", ""))
                wr = float(row["dram__bytes_write.sum"].replace(",", ""))
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                t = rd * scale[units["dram__bytes_read.sum"]] + wr * scale[units["dram__bytes_write.sum"]]
                dur = float(row["gpu__time_duration.sum"].replace(",", ""))
                dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units["gpu__time_duration.sum"]]
                print(f"  {'traffic (read+write)':34s} {t:18.0f} bytes -> {t / dur_s / 1e9:.0f} GB/s under ncu (cold, serialised)")
            except Exception as e:  # noqa: BLE001
                print(f"  traffic: n/a ({e})")
            stalls = []
            for k, v in row.items():
                if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(v.replace(",", "")), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("  top stalls (warps stalled per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))
            print()


if __name__ == "__main__":
    main()
