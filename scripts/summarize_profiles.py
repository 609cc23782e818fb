#!/usr/bin/env python
"""Turn the raw ncu artefacts of a gpurun call (gpurun_out/launches.csv, gpurun_out/prof_*.ncu-rep) into the small
text summaries committed under profiles/ (named per round).  Runs in the build container (ncu -i needs no GPU).

usage: python scripts/summarize_profiles.py r01 [tag]
"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
    # FP32-pipe evidence for the Chamfer kernel: pipe utilisation and issue-slot use
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_active.avg", "smsp__cycles_active.avg",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_tmem.sum",
]


def launches(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("ldt::", "").replace("void ", "")
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        seq.append((name, v))
    return seq


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    rnd = sys.argv[1]
    tag = ("_" + sys.argv[2]) if len(sys.argv) > 2 else ""
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    lp = os.path.join(OUT, "launches.csv")
    if os.path.exists(lp):
        seq = launches(lp)
        idx = [i for i, (n, _) in enumerate(seq) if n.startswith("select_row")]
        out = io.StringIO()
        out.write("# ncu launch list (--metrics gpu__time_duration.sum --clock-control none) of\n"
                  "#   python bench.py --sde-steps 2 --steps 1 --warmup 1 --no-cpu-baseline --cd-clouds 16\n"
                  "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n")
        if len(idx) >= 2:
            a, b = idx[-2], idx[-1]
            agg = collections.OrderedDict()
            for n, v in seq[a:b]:
                x = agg.setdefault(n, [0, 0.0])
                x[0] += 1
                x[1] += v
            tot = sum(x[1] for x in agg.values())
            out.write(f"\n## one reverse-SDE step (graph replay): {b - a} launches, {tot:.1f} us of kernel time\n")
            out.write(f"{'kernel':58s} {'n':>4s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}\n")
            for k, x in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                out.write(f"{k[:58]:58s} {x[0]:4d} {x[1]:10.1f} {x[1] / x[0]:9.2f} {x[1] / tot:6.3f}\n")
        agg = collections.OrderedDict()
        for n, v in seq:
            x = agg.setdefault(n, [0, 0.0])
            x[0] += 1
            x[1] += v
        tot = sum(x[1] for x in agg.values())
        out.write(f"\n## whole command: {len(seq)} launches captured, {tot:.1f} us\n")
        for k, x in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            out.write(f"{k[:58]:58s} {x[0]:4d} {x[1]:10.1f} {x[1] / x[0]:9.2f} {x[1] / tot:6.3f}\n")
        open(os.path.join(ROOT, "profiles", f"{rnd}_launches{tag}.txt"), "w").write(out.getvalue())
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        out = io.StringIO()
        out.write(f"# ncu --set full --clock-control none, kernel regex {name}; one column per captured launch\n")
        ki = h.index("Kernel Name")
        out.write("kernel: " + " | ".join(re.sub(r"\(.*", "", x[ki]) for x in rows[2:]) + "\n")
        for m in METRICS:
            if m in h:
                i = h.index(m)
                out.write(f"{m} [{units[i]}]: " + " | ".join(x[i] for x in rows[2:]) + "\n")
        open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_{name}{tag}.txt"), "w").write(out.getvalue())
        # per-launch DRAM traffic (bytes) for bench.py's roofline.traffic
        try:
            ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            per = [float(x[ir]) * scale[units[ir]] + float(x[iw]) * scale[units[iw]] for x in rows[2:]]
            traffic[name] = {"launches": len(per), "dram_bytes_per_launch": sum(per) / len(per),
                             "kernels": [re.sub(r"\(.*", "", x[ki]) for x in rows[2:]]}
        except Exception:
            pass
    if traffic:
        import json
        json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{rnd}_traffic{tag}.json"), "w"), indent=1)
    print("\n".join(sorted(os.listdir(os.path.join(ROOT, "profiles")))))


if __name__ == "__main__":
    main()
