"""Turn gpurun_out/ ncu artefacts into small tracked summaries under profiles/.

    python tools/summarize_profiles.py r1      # reads gpurun_out/launches.csv, gpurun_out/*.ncu-rep
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list (ncu --metrics gpu__time_duration.sum) ----
path = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    starts = [i for i, r in enumerate(data) if "fps_cluster_kernel" in r[ki] or "fps_cta_kernel<8" in r[ki]]
    step = data[starts[-2]:starts[-1]] if len(starts) >= 2 else data
    # the 512 MB L2-flush memset bench.py issues between steps is outside its timed region: drop it
    step = [r for r in step if not ("FillFunctor<unsigned char>" in r[ki] and "262144" in r[gi])]
    lines, tot, ours = [], 0.0, 0.0
    agg = collections.OrderedDict()
    for r in step:
        us = float(r[vi].replace(",", "")) / 1000.0
        tot += us
        name = r[ki].replace("cpfn::<unnamed>::", "").replace("void ", "")[:90]
        if "cpfn::" in r[ki]:
            ours += us
        lines.append("| %8.1f | %s | %s | %s |" % (us, r[gi], r[bi], name))
        a = agg.setdefault(name.split("(")[0], [0, 0.0]); a[0] += 1; a[1] += us
    with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
        f.write("# %s: every kernel of ONE bench step (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % tag)
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv python bench.py --steps 2 --warmup 1`\n"
                "(per-launch times are cold-cache and serialised: compare SHARES).  One step = %d launches, %.1f us total, "
                "%.1f us (%.0f %%) in this library's kernels.\n\n" % (len(step), tot, ours, 100 * ours / tot))
        f.write("## share by kernel\n\n| kernel | launches | us | share |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.1f | %.1f %% |\n" % (k, n, t, 100 * t / tot))
        f.write("\n## launch order\n\n| us | grid | block | kernel |\n|---|---|---|---|\n" + "\n".join(lines) + "\n")
    print("wrote", os.path.join(out_dir, "%s_launches.md" % tag))

# ---- full captures ----
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "*.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    h = rows[0]
    idx = {n: i for i, n in enumerate(h)}
    name = os.path.basename(rep)[:-8]
    with open(os.path.join(out_dir, "%s_ncu_%s.md" % (tag, name)), "w") as f:
        f.write("# %s: ncu --set full --clock-control none --import-source on (%s)\n\n" % (tag, os.path.basename(rep)))
        for r in rows[2:]:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[idx["Kernel Name"]][:100])
            for w in WANT:
                if w in idx:
                    f.write("| %s | %s | %s |\n" % (w, r[idx[w]], rows[1][idx[w]]))
            stalls = []
            for n, i in idx.items():
                if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(r[i].replace(",", "")), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            f.write("\nwarp stall reasons (warps per issue-active cycle): " +
                    ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)[:6]) + "\n\n")
    print("wrote", os.path.join(out_dir, "%s_ncu_%s.md" % (tag, name)))
    tpath = os.path.join(out_dir, "%s_traffic.json" % tag)
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * unit[rows[1][idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * unit[rows[1][idx["dram__bytes_write.sum"]]]
        except (KeyError, ValueError):
            continue
        kname = r[idx["Kernel Name"]].split("(")[0].split("::")[-1]
        key = "%s grid=%s" % (kname, r[idx["launch__grid_size"]])
        traffic[key] = {"dram_bytes": rd + wr, "us": float(r[idx["gpu__time_duration.sum"]].replace(",", ""))}
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
