"""Round-2 evidence -> profiles/: run after tools/gpu_r2_final.sh has merged its files into gpurun_out/.
  gpurun_out/r2_launches.csv        -> profiles/r2_launches.md      (every launch of one eager step + shares)
  gpurun_out/r2_step_final.ncu-rep  -> profiles/r2_ncu_step.md, profiles/r2_traffic.json   (`ncu --set full`)
  gpurun_out/r2_bench_*.json, kernel_roofline.json, r2_seg.md, r2_chain_profile.txt -> profiles/ (copied)"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"^(void )?(cpfn::)?(<?unnamed>::|\(anonymous namespace\)::)*", "", name.strip())
    name = re.sub(r"cpfn::<unnamed>::", "", name)
    return re.sub(r"\(.*$", "", name)


def num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else float("nan")


def launches():
    path = os.path.join(GO, "r2_launches.csv")
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[ix["ID"]], {"kernel": short(r[ix["Kernel Name"]]), "grid": r[ix["Grid Size"]].replace(" ", ""),
                                          "block": r[ix["Block Size"]].replace(" ", "")})
        unit, val = r[ix["Metric Unit"]], num(r[ix["Metric Value"]])
        name = r[ix["Metric Name"]]
        if name == "gpu__time_duration.sum":
            val = val / 1e3 if unit in ("ns", "nsecond") else (val * 1e3 if unit in ("ms", "msecond") else val)
        if name.startswith("dram__bytes"):
            val *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d[name] = val
    L = list(per.values())
    tot = sum(d["gpu__time_duration.sum"] for d in L)
    chains = [d for d in L if d["kernel"].startswith("mlp_chain")]
    out = ["# r2: every kernel of ONE GlobalSPFN step (B = 16 x 8192 points, K = 28, forward + fit)", "",
           "Command: `ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,smsp__issue_active...,dram__bytes_* "
           "--clock-control none --launch-skip 143 --launch-count 23 --csv python tools/profile_step.py` (second of two eager steps; "
           "per-launch times are cold-cache and serialised: compare SHARES).",
           "One step = %d launches, %.1f us, all of them this library's kernels (round 1: 25 launches, 1063 us, one ATen dropout "
           "kernel; round 2 before the chain work: 23 launches, 1013.8 us).  The %d MLP-chain launches: %.1f us (round 1: 502-530, "
           "start of round 2: 511.7)." % (len(L), tot, len(chains), sum(d["gpu__time_duration.sum"] for d in chains)), "",
           "## in launch order", "",
           "| # | kernel | grid | block | us | share | tensor pipe active % | issue active % | DRAM MB (read + write) |", "|---|---|---|---|---|---|---|---|---|"]
    for i, d in enumerate(L):
        t = d["gpu__time_duration.sum"]
        out.append("| %d | %s | %s | %s | %.1f | %.1f %% | %.1f | %.1f | %.2f |" % (
            i, d["kernel"], d["grid"], d["block"], t, 100 * t / tot,
            d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", float("nan")),
            d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", float("nan")),
            (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6))
    agg = collections.OrderedDict()
    for d in L:
        a = agg.setdefault(d["kernel"], [0, 0.0])
        a[0] += 1
        a[1] += d["gpu__time_duration.sum"]
    out += ["", "## share by kernel", "", "| kernel | launches | us | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.1f | %.1f %% |" % (k, n, t, 100 * t / tot))
    open(os.path.join(PR, "r2_launches.md"), "w").write("\n".join(out) + "\n")
    print("r2_launches.md:", len(L), "launches", round(tot, 1), "us")


def full_capture():
    rep = os.path.join(GO, "r2_step_final.ncu-rep")
    if not os.path.exists(rep):
        print("no full capture")
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}

    def get(r, key, scale=None):
        if key not in ix:
            return float("nan")
        v, u = num(r[ix[key]]), units[ix[key]]
        if scale == "us":
            return v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        if scale == "bytes":
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return v

    out = ["# r2: `ncu --set full --clock-control none --import-source on` of ONE eager GlobalSPFN step "
           "(B = 16 x 8192 points, K = 28, forward + fit)", "",
           "Command: `ncu --set full --clock-control none --import-source on --launch-skip 143 --launch-count 23 -o "
           "gpurun_out/r2_step_final python tools/profile_step.py` (second of two eager steps; kernels are serialised and replayed by "
           "ncu, so durations are cold-cache: compare shares).  End-of-round-2 code.", "",
           "| kernel | us | grid | regs | DRAM read MB | DRAM write MB | tensor pipe % | issue active % | warps active % | L1 hit % | L2 hit % | warp instructions |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic, tot, chains = {}, 0.0, 0.0
    for r in rows[2:]:
        name = short(r[ix["Kernel Name"]])
        us = get(r, "gpu__time_duration.sum", "us")
        rd, wr = get(r, "dram__bytes_read.sum", "bytes"), get(r, "dram__bytes_write.sum", "bytes")
        tot += us
        if name.startswith("mlp_chain"):
            chains += us
        out.append("| %s | %.1f | %s | %d | %.2f | %.2f | %.1f | %.1f | %.1f | %.1f | %.1f | %.2f M |" % (
            name, us, r[ix["Grid Size"]].replace(" ", ""), int(get(r, "launch__registers_per_thread")), rd / 1e6, wr / 1e6,
            get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
            if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in ix
            else get(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
            get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            get(r, "l1tex__t_sector_hit_rate.pct"), get(r, "lts__t_sector_hit_rate.pct"), get(r, "smsp__inst_executed.sum") / 1e6))
        key = "%s grid=%s" % (name, r[ix["Grid Size"]].replace(" ", ""))
        k, n = key, 2
        while k in traffic:
            k, n = "%s #%d" % (key, n), n + 1
        traffic[k] = {"dram_bytes": rd + wr, "us": round(us, 3)}
    out += ["", "Sum of the %d launches: %.1f us.  The MLP chains take %.1f us of it." % (len(rows) - 2, tot, chains)]
    open(os.path.join(PR, "r2_ncu_step.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(os.path.join(PR, "r2_traffic.json"), "w"), indent=1, sort_keys=True)
    print("r2_ncu_step.md:", round(tot, 1), "us, chains", round(chains, 1))


def copies():
    for src, dst in (("r2_bench_n1.json", "r2_bench_n1.json"), ("r2_bench_reference_arm.json", "r2_bench_reference_arm.json"),
                     ("r2_bench_n2.json", "r2_bench_n2.json"), ("r2_bench_n4.json", "r2_bench_n4.json"),
                     ("r2_bench_n8.json", "r2_bench_n8.json"), ("kernel_roofline.json", "r2_kernel_roofline.json"),
                     ("r2_seg.md", "r2_seg.md"), ("r2_chain_profile.txt", "r2_chain_profile.txt")):
        s = os.path.join(GO, src)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(PR, dst))
            print("copied", dst)


if __name__ == "__main__":
    launches()
    full_capture()
    copies()
