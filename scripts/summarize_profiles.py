#!/usr/bin/env python
"""Turn the ncu outputs under gpurun_out/ into the small text/CSV summaries committed under profiles/.
usage: python scripts/summarize_profiles.py <tag>      (e.g. r01a)"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launch_summary(csv_path: Path, out_path: Path, note: str):
    lines = [l for l in csv_path.read_text().splitlines() if l and not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r["Metric Unit"], v)
        agg.setdefault(r["Kernel Name"], []).append(v)
    total = sum(sum(v) for v in agg.values())
    with out_path.open("w") as f:
        f.write(f"# {note}\n# source: {csv_path.name} ({len(rows)} launches, gpu__time_duration.sum, ncu --clock-control none; serialised, cold-cache:\n")
        f.write("# compare SHARES, not absolutes)\n")
        f.write(f"{'total_us':>12} {'share':>7} {'n':>6} {'avg_us':>9} {'min_us':>9} {'max_us':>9}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{sum(v):12.1f} {sum(v) / total:7.3f} {len(v):6d} {sum(v) / len(v):9.2f} {min(v):9.2f} {max(v):9.2f}  {k[:150]}\n")


def full_summary(rep: Path, out_path: Path, note: str):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with out_path.open("w") as f:
        f.write(f"# {note}\n# source: {rep.name} (ncu --set full --clock-control none)\n")
        for r in rows[2:]:
            f.write(f"\nkernel {r[hdr.index('Kernel Name')][:120]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            for i, h in enumerate(hdr):
                if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h):
                    f.write(f"  {h:88s} {units[i]:16s} {r[i]}\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    PROF.mkdir(exist_ok=True)
    for name, note in (("launches_phases", "one launch per phase (AVN_LAUNCH_MODE=phases), 100k-cube stack"),
                       ("launches_mega", "default mode: one persistent megakernel per step, 100k-cube stack")):
        p = OUT / f"{name}.csv"
        if p.exists():
            launch_summary(p, PROF / f"{tag}_{name}_summary.txt", note)
    for name, note in (("mega_full", "step_megakernel<float>, 100k-cube stack, one step"),
                       ("solve_full", "phase_kernel<float, OP_SOLVE_BIAS> (solver-iteration kernel), 100k-cube stack, 3 launches"),
                       ("sweep_full", "broad-phase sweep kernel, 100k-cube stack")):
        p = OUT / f"{name}.ncu-rep"
        if p.exists():
            full_summary(p, PROF / f"{tag}_{name}_ncu.txt", note)
    for name in ("bench_100k.log", "bench_100k_phases.log", "pytest_gpu.log", "gpu.txt"):
        p = OUT / name
        if p.exists():
            (PROF / f"{tag}_{name}").write_text(p.read_text())
