"""Turns the ncu reports / launch list a tools/gpu_round.sh visit left in gpurun_out/ into the small tracked files under
profiles/ (run here, on the CPU box):  python tools/summarize_profiles.py <tag>
  profiles/<tag>_launches.csv        per-launch device times of one forward (ncu, serialised)
  profiles/<tag>_launch_shares.md    the same, aggregated per kernel with shares of the forward
  profiles/<tag>_<name>_raw.csv      key raw metrics of the captured kernels (dram bytes, pipe activity, stalls)
  profiles/traffic.json              dram bytes per launch of the dominant kernels (read by bench.py -> roofline.traffic)
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second"]

lp = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(dst, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(lp)) if len(r) > 10 and r[0] != "ID"]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split("(")[0], []).append(float(r[-1]))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(dst, f"{tag}_launch_shares.md"), "w") as f:
        f.write(f"# {tag}: launches of one config-2 forward + sampler step (ncu gpu__time_duration, serialised, cold cache)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.2f} | {sum(v)/tot*100:.1f} % |\n")
        f.write(f"\ntotal {tot/1e3:.1f} us over {len(rows)} launches\n")

traffic = {}
for name in ("attn", "gemm", "pointwise", "conv"):
    rp = os.path.join(src, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rp):
        continue
    out = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    extra = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    cols = [h for h in KEEP + extra if h in hdr]
    idx = [hdr.index(h) for h in cols]
    with open(os.path.join(dst, f"{tag}_{name}_raw.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])
    def val(r, h):
        return float(r[hdr.index(h)]) if h in hdr and r[hdr.index(h)] not in ("", "n/a") else None
    def to_bytes(r, h):
        v = val(r, h)
        if v is None:
            return None
        u = units[hdr.index(h)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    for r in rows[2:]:
        kn = r[hdr.index("Kernel Name")].split("(")[0] if "Kernel Name" in hdr else name
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        if rd is not None and wr is not None:
            traffic.setdefault(kn, []).append({"dram_bytes": rd + wr, "us": val(r, "gpu__time_duration.sum"), "source": f"profiles/{tag}_{name}_raw.csv"})
if traffic:
    json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
bl = os.path.join(src, f"{tag}_bench.log")
if os.path.exists(bl):
    lines = [x for x in open(bl) if x.startswith("{")]
    if lines:
        open(os.path.join(dst, f"{tag}_bench.json"), "w").write(lines[-1])
print("profiles/ updated:", sorted(x for x in os.listdir(dst) if x.startswith(tag) or x == "traffic.json"))
