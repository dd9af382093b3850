#!/usr/bin/env python3
"""Turns gpurun_out/*.ncu-rep and launch lists into small tracked summaries under profiles/."""
import csv, collections, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    out = []
    for r in rows[2:]:
        d = {h: (rows[1][i], r[i]) for i, h in enumerate(hdr)}
        out.append(d)
    return out


def summarize_rep(rep, name, note):
    launches = ncu_raw(rep)
    with open(os.path.join(OUT, name), "w") as f:
        f.write(f"# {name}\n# source: {os.path.basename(rep)} (ncu --set full --clock-control none --import-source on)\n# {note}\n")
        for k, d in enumerate(launches):
            f.write(f"\n[launch {k}] {d.get('Kernel Name', ('', '?'))[1]}\n")
            for m in WANT:
                if m in d:
                    f.write(f"  {m:90s} {d[m][1]:>18s} {d[m][0]}\n")
            try:
                t = float(d["gpu__time_duration.sum"][1].replace(",", ""))
                tu = d["gpu__time_duration.sum"][0]
                scale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[tu]

                def gb(key):
                    v, u = float(d[key][1].replace(",", "")), d[key][0]
                    return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}[u]
                tr = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
                f.write(f"  => dram traffic {tr:.3f} GB in {t * scale * 1e3:.3f} ms = {tr / (t * scale):.1f} GB/s\n")
            except Exception as e:
                f.write(f"  (traffic summary failed: {e})\n")


def summarize_launches(csvfile, name, note):
    rows = [r for r in csv.reader(open(csvfile)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in rows[1:]:
        try:
            seq.append((r[ki].split("(")[0][:70], float(r[vi].replace(",", ""))))
        except ValueError:
            pass
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("k_aabb")]
    a, b = (starts[0], starts[1]) if len(starts) > 1 else (0, len(seq))
    agg = collections.OrderedDict()
    for n, v in seq[a:b]:
        e = agg.setdefault(n, [0.0, 0]); e[0] += v; e[1] += 1
    tot = sum(v[0] for v in agg.values())
    with open(os.path.join(OUT, name), "w") as f:
        f.write(f"# {name}\n# source: {os.path.basename(csvfile)} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised:\n"
                f"# compare SHARES).  One reconstruct step = launches between two k_aabb.  {note}\n")
        f.write(f"{'kernel':72s} {'ms':>10s} {'launches':>9s} {'share':>7s}\n")
        for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
            f.write(f"{n:72s} {v / 1e6:10.3f} {c:9d} {100 * v / tot:6.1f}%\n")
        f.write(f"{'total':72s} {tot / 1e6:10.3f}\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    jobs = [("prof_levelset_r1a.ncu-rep", "r1a_levelset_exact_everywhere_ncu.txt", "round 1, first version: every grid point evaluated exactly (2 M-particle dam break)"),
            ("prof_levelset_r1b.ncu-rep", "r1b_levelset_certify_ncu.txt", "round 1: certification pass added (4 M particles); launch 0 = certify, launch 1 = fix-up"),
            ("prof_levelset_r1c.ncu-rep", "r1c_levelset_certify_opt_ncu.txt", "round 1: prologue/staging/FMA-certification optimised (4 M particles)"),
            ("prof_levelset_r1d.ncu-rep", "r1d_levelset_worklist_ext_ncu.txt", "round 1: work list of non-empty bricks + extension bricks + two-ring certification (4 M particles; before the cubic lower bound)")]
    for rep, name, note in jobs:
        if os.path.exists(os.path.join(g, rep)):
            summarize_rep(os.path.join(g, rep), name, note)
    for c, name, note in [("launches_r1a.csv", "r1a_launch_list.txt", "2 M particles, exact-everywhere version"),
                          ("launches_r1b.csv", "r1b_launch_list.txt", "10 M particles, certification version"),
                          ("launches_r1c.csv", "r1c_launch_list_50M.txt", "`python bench.py --steps 1 --warmup 0` (50 M particles), plane-indexed MC passes"),
                          ("launches_r1d.csv", "r1d_launch_list_50M.txt", "`python bench.py --steps 1 --warmup 0` (50 M particles), brick-list MC passes, before two-ring certification / work list / extension bricks")]:
        if os.path.exists(os.path.join(g, c)):
            summarize_launches(os.path.join(g, c), name, note)
    print(os.listdir(OUT))
