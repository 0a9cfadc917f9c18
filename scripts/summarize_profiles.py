#!/usr/bin/env python
"""Turn the ncu outputs scripts/gpu_profile.sh left under gpurun_out/ into the small text summaries committed under profiles/, and
regenerate profiles/traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py for `roofline.traffic`) FROM THE SAME CAPTURE.
usage: python scripts/summarize_profiles.py <tag> [scene ...]      (e.g. r02b stack100k spheres1m)"""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"]
GB = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


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


def full_summary(raw_csv: Path, out_path: Path, note: str):
    """raw_csv = `ncu -i X.ncu-rep --page raw --csv`; returns (kernel name, dram bytes per launch) of the first row"""
    rows = list(csv.reader(raw_csv.read_text().splitlines()))
    hdr, units = rows[0], rows[1]
    first = None
    with out_path.open("w") as f:
        f.write(f"# {note}\n# source: {raw_csv.name} (ncu --set full --clock-control none --import-source on, --page raw)\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"\nkernel {name[:160]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n")
            vals = {}
            for i, h in enumerate(hdr):
                if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h):
                    f.write(f"  {h:88s} {units[i]:16s} {r[i]}\n")
                    vals[h] = (r[i], units[i])
            if first is None and "dram__bytes_read.sum" in vals:
                tot = sum(float(vals[k][0].replace(",", "")) * GB.get(vals[k][1], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                first = (name, int(tot))
    return first


if __name__ == "__main__":
    tag = sys.argv[1]
    scenes = sys.argv[2:] or ["stack100k", "spheres1m"]
    PROF.mkdir(exist_ok=True)
    tfile = PROF / "traffic.json"
    traffic = json.loads(tfile.read_text()) if tfile.exists() else {}
    traffic["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from one `ncu --set full` capture of the tree "
                        "named in `capture`; written by scripts/summarize_profiles.py together with the text summary it cites")
    for scene in scenes:
        p = OUT / f"{tag}_launches_{scene}.csv"
        if p.exists():
            launch_summary(p, PROF / f"{tag}_launches_{scene}_summary.txt", f"default mode: one persistent megakernel per step + the broad phase, {scene}")
        p = OUT / f"{tag}_mega_{scene}_raw.csv"
        if p.exists():
            got = full_summary(p, PROF / f"{tag}_mega_{scene}_ncu.txt", f"step_megakernel, {scene}, one step")
            if got:
                traffic[scene] = {"kernel": got[0][:120], "dram_bytes_per_launch": got[1], "capture": f"{tag} (profiles/{tag}_mega_{scene}_ncu.txt)"}
        p = OUT / f"{tag}_mega_{scene}_details.txt"
        if p.exists():
            (PROF / f"{tag}_mega_{scene}_ncu_details.txt").write_text(p.read_text())
    tfile.write_text(json.dumps(traffic, indent=1))
    print(json.dumps(traffic, indent=1))
