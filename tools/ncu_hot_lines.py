"""Hot source lines of one kernel of an .ncu-rep (captured with --import-source on, built with -lineinfo):
    python tools/ncu_hot_lines.py REPORT KERNEL_REGEX [min_share]
Prints, per CUDA source line, the stall samples, the executed warp instructions and the dominant stall reasons."""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Function Name":
            cur = {"name": r[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif r and r[0] == "Line No" and cur is not None:
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and r and r[0].strip().isdigit():
            cur["rows"].append(r)
    for b in blocks[:1]:
        h = b["hdr"]
        si, ie = h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        lines = {}
        for r in b["rows"]:
            if len(r) <= si or not r[si].isdigit():
                continue
            ln = int(r[0])
            d = lines.setdefault(ln, {"src": r[1], "s": 0, "e": 0, "st": {}})
            d["s"] += int(r[si]); d["e"] += int(r[ie]) if r[ie].isdigit() else 0
            for i, c in stall_cols:
                if r[i].isdigit():
                    d["st"][c] = d["st"].get(c, 0) + int(r[i])
        tot = sum(d["s"] for d in lines.values()); tote = sum(d["e"] for d in lines.values())
        print("%s\n total samples %d, warp instructions %d" % (b["name"], tot, tote))
        for ln in sorted(lines):
            d = lines[ln]
            if d["s"] >= tot * min_share or d["e"] >= tote * min_share:
                top = sorted(d["st"].items(), key=lambda kv: -kv[1])[:3]
                print("%5d %5.1f%% smp %5.1f%% ins | %-90s | %s" % (ln, 100.0 * d["s"] / tot, 100.0 * d["e"] / max(1, tote),
                      d["src"].strip()[:90], ", ".join("%s %d" % (k[6:], v) for k, v in top if v)))


if __name__ == "__main__":
    main()
