import csv, io, subprocess, sys
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    for r in rows[2:3]:
        d = dict(zip(h, r))
        st = []
        for k, v in d.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                try: st.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except Exception: pass
        print(path.split("/")[-1], d["Kernel Name"][:40], "| stalls (warps per issue):", ", ".join("%s %.2f" % (k, v) for v, k in sorted(st, reverse=True)[:7]))
