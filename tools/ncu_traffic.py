"""Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel
in a set of .ncu-rep files -> JSON keyed by the (namespace-stripped) kernel name.
usage: python tools/ncu_traffic.py out.json rep1 [rep2 ...]
       python tools/ncu_traffic.py --calls profiles/ncu_traffic.json rep1 [rep2 ...]
--calls writes the table bench.py reads: mean DRAM bytes per launch keyed by the C-ABI call
the kernel belongs to, plus the digest of the CUDA sources (bench.source_digest), so that
bench.py only uses it while the library is built from those sources."""
import os
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


# C-ABI call -> kernels whose per-launch traffic adds up to one call (cfg3, M = 1)
CALL_KERNELS = {
    "chb_push_depose_push_index": ["depose_kernel<1, 1, 2, 32>"],
    "chb_gather_push": ["gather_push_kernel<1, 0>"],
    "chb_depose_scalar": ["depose_kernel<1, 0, 0, 128>"],
    "chb_sort_scatter_stable": ["sort_scatter_kernel", "sort_fixup_kernel"],
    "chb_psatd_advance": ["psatd_kernel"],
    "chb_fft_x_batched": ["fft_pow2_kernel<12>"],
    "chb_fft_damp_x_batched": ["fft_damp_kernel<12>"],
    "chb_dht_batched": ["dht_gemm_wide_kernel<7, 8>"],
}


def main():
    calls = sys.argv[1] == "--calls"
    if calls:
        sys.argv.pop(1)
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
    if calls:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench
        norm = {k.replace("void ", "").strip(): v for k, v in res.items()}
        table = {}
        for call, kernels in CALL_KERNELS.items():
            if all(k in norm for k in kernels):
                tot = sum(sum(norm[k]["dram_bytes_per_launch"]) / len(norm[k]["dram_bytes_per_launch"])
                          for k in kernels)
                table[call] = {"dram_bytes_per_launch": tot, "kernel": " + ".join(kernels)}
        res = {"source_digest": bench.source_digest(), "calls": table, "kernels": res}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    if calls:
        for k, v in res["calls"].items():
            print(k, "%.1f MB" % (v["dram_bytes_per_launch"] / 1e6), v["kernel"])
        return
    for k, v in res.items():
        print(k, ["%.1f MB" % (b / 1e6) for b in v["dram_bytes_per_launch"]])


if __name__ == "__main__":
    main()
