#!/usr/bin/env python3
"""Executed warp-instructions of a profiled kernel, grouped by the device function they come from.

Input: an .ncu-rep captured with `--set full --import-source on` (kernels built with -lineinfo) and the git revision the
capture was made from.  `ncu --page source --print-source cuda,sass` gives executed instructions per source line; the line ->
function map is recovered from the sources of that revision.

    python tools/source_breakdown.py gpurun_out/prof_levelset_r1d.ncu-rep 9d5864e > profiles/r1d_levelset_source_breakdown.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

FUNC = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__global__|__device__|__host__)[^;{]*?\b(\w+)\s*\(")


def function_map(rev, path):
    text = subprocess.run(["git", "show", f"{rev}:{path}"], capture_output=True, text=True, check=True).stdout.splitlines()
    owner, cur, pending = {}, None, ""
    for no, line in enumerate(text, 1):
        joined = (pending + " " + line).strip()
        m = FUNC.match(joined) or re.match(r"^(k_\w+|ss_\w+)\s*\(", line)
        if m:
            cur = m.group(1)
        pending = line if line.startswith(("template", "__global__", "__device__")) and "(" not in line else ""
        owner[no] = cur
    return owner


def main():
    rep, rev = sys.argv[1], sys.argv[2]
    kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []     # optional: one kernel of a multi-kernel report
    nth = ["--launch-skip", sys.argv[4], "--launch-count", "1"] if len(sys.argv) > 4 else []
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kern + nth,
                         capture_output=True, text=True, check=True).stdout
    maps = {}
    agg, smp, src = collections.Counter(), collections.Counter(), {}
    cur, col, col_s = None, None, None
    kernel = None
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1]
        elif len(r) >= 2 and r[0] == "Function Name":
            kernel = kernel or r[1]
        elif r and r[0] == "Line No":
            col = r.index("Instructions Executed"); col_s = r.index("# Samples")
        elif col is not None and r and r[0].isdigit():
            def num(v):
                try:
                    return int(v)
                except ValueError:
                    return 0
            agg[(cur, int(r[0]))] += num(r[col]); smp[(cur, int(r[0]))] += num(r[col_s]); src[(cur, int(r[0]))] = r[1]
    total = sum(agg.values())
    per_func, per_func_s = collections.Counter(), collections.Counter()
    total_s = max(sum(smp.values()), 1)
    for (path, no), v in agg.items():
        name = path.split("/")[-1]
        if "splashsurf_b200/csrc/" in path:
            rel = path[path.index("splashsurf_b200/csrc/"):]
            if rel not in maps:
                maps[rel] = function_map(rev, rel)
            key = f"{name}: {maps[rel].get(no)}"
        else:
            key = f"{name} (toolkit header)"
        per_func[key] += v; per_func_s[key] += smp[(path, no)]
    print(f"# {kernel}\n# report {rep.split('/')[-1]}, sources of revision {rev}; total {total:,} executed warp-instructions\n")
    print("## by device function: share of executed instructions | share of warp-state samples (time)")
    for k, v in per_func.most_common():
        print(f"{100.0 * v / total:6.2f} %  {100.0 * per_func_s[k] / total_s:6.2f} %  {v:>14,}  {k}")
    print("\n## hottest source lines (instructions | samples)")
    for (path, no), v in agg.most_common(40):
        print(f"{100.0 * v / total:6.2f} %  {100.0 * smp[(path, no)] / total_s:6.2f} %  {path.split('/')[-1]}:{no:<5} {src[(path, no)].strip()[:110]}")


if __name__ == "__main__":
    main()
