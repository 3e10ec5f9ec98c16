"""Print duration, DRAM traffic, occupancy, pipes and the top stall reasons of each launch
in .ncu-rep files.  usage: python tools/ncu_top.py rep [rep ...]"""
import csv, io, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3,
        "s": 1e6}
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    u = dict(zip(h, units))
    for r in rows[2:]:
        d = dict(zip(h, r))

        def f(k, scale=True):
            try:
                v = float(d[k].replace(",", ""))
            except Exception:
                return float("nan")
            return v * UNIT.get(u.get(k, ""), 1.0) if scale else v
        print("== %s | %s" % (path.split("/")[-1], d["Kernel Name"][:70]))
        print("   dur %.1f us  dram R %.1f MB W %.1f MB  dram%% %.1f  regs %s  grid %s block %s  waves %s  occ(warps act%%) %.1f  issue%% %.1f" % (
            f("gpu__time_duration.sum"), f("dram__bytes_read.sum") / 1e6, f("dram__bytes_write.sum") / 1e6,
            f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False), d["launch__registers_per_thread"],
            d["launch__grid_size"], d["launch__block_size"], d["launch__waves_per_multiprocessor"],
            f("sm__warps_active.avg.pct_of_peak_sustained_active", False),
            f("smsp__issue_active.avg.pct_of_peak_sustained_active", False)))
        print("   L1 hit %.1f L2 hit %.1f | fp64 pipe %.1f dmma pipe %.1f lsu %.1f" % (
            f("l1tex__t_sector_hit_rate.pct", False), f("lts__t_sector_hit_rate.pct", False),
            f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False),
            f("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", False),
            f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False)))
        st = []
        pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        for k, v in d.items():
            if k.startswith(pre) and k.endswith(suf):
                try:
                    st.append((float(v), k[len(pre):-len(suf)]))
                except Exception:
                    pass
        print("   stalls (warps per issue):", ", ".join("%s %.2f" % (k, v) for v, k in sorted(st, reverse=True)[:6]))
