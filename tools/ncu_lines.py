"""Per-source-line stall samples from an .ncu-rep captured with --import-source on
(kernels built with -lineinfo).  usage: python tools/ncu_lines.py rep [kernel-substr] [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source",
                          "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, kern, hdr = "", "", None
    seen = set()
    agg = defaultdict(lambda: defaultdict(float))
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            kern = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or want not in kern:
            continue
        d = dict(zip(hdr, r))
        key = (kern, fname, r[0], r[1].strip()[:90])
        a = agg[key]
        try:
            a["samples"] += float(d.get("# Samples", 0) or 0)
            a["inst"] += float(d.get("Instructions Executed", 0) or 0)
        except ValueError:
            continue
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    a[k] += float(v or 0)
                except ValueError:
                    pass
    by_kernel = defaultdict(list)
    for (kern, fname, line, src), a in agg.items():
        by_kernel[kern].append((a["samples"], fname, line, src, a))
    for kern, lst in by_kernel.items():
        tot = sum(x[0] for x in lst) or 1.0
        print("== %s  (samples %d)" % (kern[:100], tot))
        for s, fname, line, src, a in sorted(lst, reverse=True)[:top]:
            st = sorted(((v, k[6:]) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:3]
            print("  %5.1f%% %s:%s  %-90s | %s | inst %d" % (
                100 * s / tot, fname, line, src, " ".join("%s %.0f" % (k, v) for v, k in st),
                a["inst"]))


if __name__ == "__main__":
    main()
