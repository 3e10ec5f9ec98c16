"""Print duration, DRAM traffic, and the top stall reasons / pipes of each launch."""
import csv, io, subprocess, sys
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    for r in rows[2:]:
        d = dict(zip(h, r))
        def f(k):
            try: return float(d[k].replace(",", ""))
            except Exception: return float("nan")
        print("== %s | %s" % (path.split("/")[-1], d["Kernel Name"][:70]))
        print("   dur %.1f us  dram R %.1f MB W %.1f MB  dram%% %.1f  regs %s  grid %s block %s  waves %s  occ(warps act%%) %.1f  issue%% %.1f" % (
            f("gpu__time_duration.sum") / (1000 if "ns" in rows[1][h.index("gpu__time_duration.sum")] else 1),
            f("dram__bytes_read.sum"), f("dram__bytes_write.sum"),
            f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), d["launch__registers_per_thread"],
            d["launch__grid_size"], d["launch__block_size"], d["launch__waves_per_multiprocessor"],
            f("sm__warps_active.avg.pct_of_peak_sustained_active"), f("smsp__issue_active.avg.pct_of_peak_sustained_active")))
        print("   units:", rows[1][h.index("dram__bytes_read.sum")], "| L1 hit %.1f L2 hit %.1f | fp64 pipe %.1f lsu %.1f" % (
            f("l1tex__t_sector_hit_rate.pct"), f("lts__t_sector_hit_rate.pct"),
            f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")))
        st = []
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("_per_warp_active.pct"):
                try: st.append((float(v), k.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")))
                except Exception: pass
        print("   stalls:", ", ".join("%s %.0f" % (k, v) for v, k in sorted(st, reverse=True)[:6]))
