"""Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel
in a set of .ncu-rep files -> JSON keyed by the (namespace-stripped) kernel name.
usage: python tools/ncu_traffic.py out.json rep1 [rep2 ...]"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    res = {}
    for path in reps:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                             text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            continue
        h, units = rows[0], rows[1]
        ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        for r in rows[2:]:
            name = r[h.index("Kernel Name")].split("(")[0].replace("chb::", "").strip()
            tot = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
            e = res.setdefault(name, {"dram_bytes_per_launch": [], "duration_us": []})
            e["dram_bytes_per_launch"].append(tot)
            d = float(r[h.index("gpu__time_duration.sum")])
            u = units[h.index("gpu__time_duration.sum")]
            e["duration_us"].append(d * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0))
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print(k, ["%.1f MB" % (b / 1e6) for b in v["dram_bytes_per_launch"]])


if __name__ == "__main__":
    main()
