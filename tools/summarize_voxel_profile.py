"""Condenses a voxelization visit (gpurun_out/<tag>_voxel_*.{json,csv,log}) into profiles/ (run on the CPU box):
    python tools/summarize_voxel_profile.py <tag>"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for suffix in ("voxel_bench.json", "voxel_launches.csv", "voxel_gpu_tests.log"):
    p = os.path.join(src, f"{tag}_{suffix}")
    if os.path.exists(p):
        shutil.copy(p, os.path.join(dst, f"{tag}_{suffix}"))
bench = json.load(open(os.path.join(src, f"{tag}_voxel_bench.json")))
lines = [l for l in open(os.path.join(src, f"{tag}_voxel_launches.csv")) if not l.startswith("==")]
launches = collections.OrderedDict()
for row in csv.DictReader(lines):
    key = (int(row["ID"]), row["Kernel Name"])
    launches.setdefault(key, {})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))


def short(name):
    import re
    m = re.search(r"(\w+_kernel)(<[^(]*>)?", name)
    if not m:
        return name.split("(")[0]
    targ = (m.group(2) or "")
    targ = re.sub(r"[\w:<> ()]*::", "", targ)  # drop namespaces inside the template argument
    return m.group(1) + (f"<{targ.strip('<>')}>" if targ else "")


# tools/profile_voxelize.py runs the fused call first, then the dense one; a call starts at voxel_insert_kernel
calls = []
for (i, name), m in launches.items():
    if short(name).startswith("voxel_insert_kernel"):
        calls.append([])
    calls[-1].append((short(name), m["gpu__time_duration.sum"] / 1e3, m["dram__bytes_read.sum"] / 1e6,
                      m["dram__bytes_write.sum"] / 1e6))
n_points = int(bench["workload"].split(",")[1].split()[0])
out = [f"# {tag}: launches of `orvb_hard_voxelize` (ncu gpu__time_duration + DRAM bytes, serialised, cold cache)", "",
       f"Workload: `tools/profile_voxelize.py` — {bench['workload']}; {bench['voxels']} voxels produced.",
       "First call = fused label vote (no dense voxel tensor), second = the reference-shaped dense output.", ""]
for title, c in zip(["fused label vote (`voxel_labels`)", "dense `[max_voxels, max_points, 4]` output"], calls):
    agg = collections.OrderedDict()
    for name, us, rd, wr in c:
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += rd; a[3] += wr
    tot = sum(a[1] for a in agg.values())
    rd = sum(a[2] for a in agg.values())
    wr = sum(a[3] for a in agg.values())
    out += [f"## {title}: {len(c)} launches, {tot:.1f} us, DRAM read {rd:.1f} MB + write {wr:.1f} MB = "
            f"{(rd + wr) * 1e6 / n_points:.0f} B per point", "",
            "| kernel | launches | total us | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {a[2]:.1f} | {a[3]:.1f} |")
    out.append("")
f, d = bench["fused_points_to_voxels"], bench["reference_shaped_voxelization"]
out += [f"Live (CUDA events, `tools/bench_voxelize.py`, warm, output / workspace allocation and table memsets included): "
        f"fused {f['ms'] * 1e3:.0f} us = {f['points_per_s'] / 1e9:.2f} G points/s "
        f"(algorithmic {f['roofline']['achieved']:.1f} GB/s = {100 * f['roofline']['frac']:.1f} % of {f['roofline']['peak']} GB/s), "
        f"dense {d['ms'] * 1e3:.0f} us = {d['points_per_s'] / 1e9:.2f} G points/s ({100 * d['roofline']['frac']:.1f} %).",
        f"Reference CPU voxelizer (`oracle/_ref`, {bench['cpu_baseline'].get('cores', '?')} core, same box): "
        f"{bench['cpu_baseline'].get('value', 0) / 1e6:.2f} M points/s."]
open(os.path.join(dst, f"{tag}_voxel_launch_shares.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
