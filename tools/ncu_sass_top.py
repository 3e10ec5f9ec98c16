"""Top SASS instructions by stall samples.  usage: python tools/ncu_sass_top.py rep kernel-substr [min_pct] [launch_index]"""
import csv, io, subprocess, sys
path, want = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern = None; hdr = None; blocks = []
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name': kern = r[1]; blocks.append((kern, [])); continue
    if r[0] == 'Address': hdr = r; continue
    if hdr and blocks: blocks[-1][1].append(dict(zip(hdr, r)))
sel = [b for b in blocks if want in b[0]]
# a kernel's launches are concatenated in one block: split at address wrap
kern, lst = sel[0]
launches = [[]]
prev = -1
for d in lst:
    a = int(d['Address'], 16)
    if a < prev: launches.append([])
    launches[-1].append(d); prev = a
lst = launches[min(which, len(launches) - 1)]
tot = sum(float(d['# Samples'] or 0) for d in lst)
print(kern[:90], 'samples', tot, 'instructions', len(lst), 'launches', len(launches))
agg = {}
for d in lst:
    for k, v in d.items():
        if k.startswith('stall_') and 'Not Issued' not in k:
            agg[k[6:]] = agg.get(k[6:], 0) + float(v or 0)
print('  totals:', ' '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for i, d in enumerate(lst):
    s = float(d['# Samples'] or 0)
    if 100 * s / tot >= minpct:
        st = sorted(((float(v or 0), k[6:]) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k), reverse=True)[:2]
        print('%5d %5.1f%% %-64s %s x%s' % (i, 100 * s / tot, d['Source'].strip()[:64], ' '.join('%s %.0f' % (k, v) for v, k in st), d['Instructions Executed']))
