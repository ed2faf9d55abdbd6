"""gpurun_out/r2_all_kernels.csv (ncu --csv of tools/ncu_per_kernel.sh) -> the per-kernel markdown table of profiles/."""
import csv
import json
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
FP64_PEAK = 36.3  # TFLOP/s, FP64 FMA probe of bench.py on the same pool
try:
    HBM_PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    HBM_PEAK = 6535.7


def main(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols = {c: i for i, c in enumerate(rows[hdr])}
    launches = defaultdict(dict)  # id -> metric -> value
    names = {}
    for r in rows[hdr + 1:]:
        if len(r) <= cols["Metric Value"]:
            continue
        lid = r[cols["ID"]]
        names[lid] = r[cols["Kernel Name"]]
        v = r[cols["Metric Value"]].replace(",", "")
        try:
            v = float(v)
        except ValueError:
            continue
        unit = r[cols["Metric Unit"]]
        m = r[cols["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> us
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        launches[lid][m] = v
    by = defaultdict(list)
    for lid, m in launches.items():
        by[names[lid]].append(m)
    print("| kernel | launches | longest launch us | executed FP64 TFLOP/s (of %.1f) | DRAM GB/s (of %.0f) | FP64 pipe %% | "
          "L1/shared data pipe %% | issue %% | L2 hit %% | regs | grid x block |" % (FP64_PEAK, HBM_PEAK))
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    order = sorted(by, key=lambda k: -sum(m.get("gpu__time_duration.sum", 0) for m in by[k]))
    for k in order:
        ms = by[k]
        top = max(ms, key=lambda m: m.get("gpu__time_duration.sum", 0))
        us = top.get("gpu__time_duration.sum", 0.0)
        g = lambda n: top.get(n, 0.0)
        flop = 2 * g("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum") + g("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum") \
            + g("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
        tf = flop / (us * 1e-6) / 1e12 if us else 0.0
        gb = (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) / (us * 1e-6) / 1e9 if us else 0.0
        name = k.replace("fheram::", "").replace("|", "\\|")
        print(f"| `{name}` | {len(ms)} | {us:.1f} | {tf:.2f} ({100 * tf / FP64_PEAK:.1f} %) | {gb:.0f} ({100 * gb / HBM_PEAK:.1f} %) | "
              f"{g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{g('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {g('lts__t_sector_hit_rate.pct'):.0f} | "
              f"{int(g('launch__registers_per_thread'))} | {int(g('launch__grid_size'))} x {int(g('launch__block_size'))} |")


if __name__ == "__main__":
    main(sys.argv[1])
