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


def write_levelset_traffic(rep, sha_file, workload, particles, fallback_rep=None):
    """profiles/levelset_traffic.json: DRAM bytes of the level-set kernels of ONE step (certification + exact pass + fix-up exact pass)
    from an `ncu --set full` capture of the bench workload, tagged with the hash of the level-set sources it was taken on
    (bench.py refuses it for any other build or workload)."""
    import json
    launches = ncu_raw(rep)
    total, parts = 0.0, []
    seen_certify = 0
    for d in launches:
        name = d.get("Kernel Name", ("", "?"))[1]
        if not any(k in name for k in ("k_certify_warp", "k_exact_warp", "k_levelset")):
            continue
        seen_certify += "k_certify_warp" in name
        if seen_certify > 1:                        # the capture ran into the next step (bench.py's instrumented step): one step only
            break

        def b(key):
            v, u = float(d[key][1].replace(",", "")), d[key][0]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        by = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
        part = {"kernel": name.split("(")[0], "grid": int(d["launch__grid_size"][1].replace(",", "")),
                "ms": float(d["gpu__time_duration.sum"][1].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[d["gpu__time_duration.sum"][0]]}
        if by != by and fallback_rep and os.path.exists(fallback_rep):
            # ncu returned NaN for this launch (its replays did not agree): bytes per CTA of the same kernel in another capture
            best = None
            for e in ncu_raw(fallback_rep):
                if e.get("Kernel Name", ("", ""))[1].split("(")[0] == part["kernel"]:
                    def bb(key, e=e):
                        v, u = float(e[key][1].replace(",", "")), e[key][0]
                        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                    g2 = int(e["launch__grid_size"][1].replace(",", ""))
                    if best is None or g2 > best[0]:
                        best = (g2, bb("dram__bytes_read.sum") + bb("dram__bytes_write.sum"))
            if best:
                by = best[1] / best[0] * part["grid"]
                part["estimated"] = f"NaN in this capture; {best[1] / best[0]:.0f} B per CTA of the same kernel in {os.path.basename(fallback_rep)} (cfg-3, {best[0]} CTAs) x this launch's grid"
        part["dram_bytes"] = by
        total += by
        parts.append(part)
    shas = open(sha_file).read().split()
    out = {"kernel": "level-set stage: k_certify_warp + k_exact_warp (exact pass, fix-up pass)", "dram_bytes_per_step": total, "launches": parts,
           "workload": workload, "particles": particles, "source_sha": shas[0], "levelset_source_sha": shas[1],
           "source": f"profiles <- {os.path.basename(rep)} (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, one step of `python bench.py --steps 1 --warmup 0`)"}
    json.dump(out, open(os.path.join(OUT, "levelset_traffic.json"), "w"), indent=1)
    return out


def sass_evidence(name):
    """Mnemonic counts that prove the sm_100a-specific paths are in the shipped binary (B200_PROFILING.md): packed FP32 (FFMA2 / FADD2 /
    FMUL2), bulk asynchronous copies (UBLKCP) and mbarrier transaction waits (SYNCS)."""
    lib = os.path.join(ROOT, "splashsurf_b200", "libsplashsurf_b200.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in txt.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            counts[cur] = collections.Counter()
        elif cur:
            for m in ("FFMA2", "FADD2", "FMUL2", "UBLKCP", "SYNCS", "FFMA ", "LDS.128", "BAR.SYNC"):
                if m in line:
                    counts[cur][m.strip()] += 1
    with open(os.path.join(OUT, name), "w") as f:
        f.write(f"# {name}: SASS mnemonic counts per kernel of splashsurf_b200/libsplashsurf_b200.so (cuobjdump -sass; sm_100a)\n")
        f.write("# FFMA2/FADD2/FMUL2 = packed FP32 pairs, UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier arrive/expect_tx/try_wait, BAR.SYNC = CTA barrier\n")
        for k, c in counts.items():
            if any(s in k for s in ("k_certify_warp", "k_exact_warp", "k_density_cells", "k_mc_", "k_fixup_flags_warp", "k_levelset", "k_densityILb0")):
                f.write(f"{k[:96]:96s} " + " ".join(f"{m}={c.get(m, 0)}" for m in ("FFMA2", "FADD2", "FMUL2", "UBLKCP", "SYNCS", "LDS.128", "BAR.SYNC")) + "\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(ROOT, "gpurun_out")
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2k"
    rep = os.path.join(g, f"prof_cfg4_{tag}.ncu-rep")
    if os.path.exists(rep):
        summarize_rep(rep, "r2_cfg4_kernels_ncu.txt", "round 2, bench workload (cfg-4, 50 M particles), one step: cell-cooperative density kernel, warp-per-brick certification, "
                      "exact pass + fix-up exact pass, fix-up sweep, marching cubes count + emit")
        print(write_levelset_traffic(rep, os.path.join(g, f"source_sha_{tag}.txt"), "cfg4", 50000400, os.path.join(g, "prof_variant2_r2i.ncu-rep")))
    for c, name, note in [(f"launches_{sys.argv[2] if len(sys.argv) > 2 else 'r2j'}.csv", "r2_launch_list_50M.txt", "`python bench.py --steps 2 --warmup 1` (cfg-4, 50 M particles), round-2 kernels"),
                          ("launches_cfg5_r2j.csv", "r2_launch_list_cfg5_overlap.txt", "cfg-5 with SUPERIMPOSED droplets (the round-2 first-run cloud, now `--workload cfg5_overlap`), 200 M particles, one GPU, "
                           "before the 1024-candidate exact variant: the 4096-candidate kernel (one warp per CTA, 2 CTAs per SM) takes 62 % of the step")]:
        if os.path.exists(os.path.join(g, c)):
            summarize_launches(os.path.join(g, c), name, note)
    sass_evidence("r2_sass_evidence.txt")
    print(sorted(os.listdir(OUT)))
