"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report exported with
   ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv
usage: python scripts/ncu_lines.py x.csv [top_n]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
def num(x):
    try: return int(x)
    except Exception: return 0
cur = None; agg = collections.Counter(); src = {}; samp = collections.Counter(); thr = collections.Counter(); hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if len(r) > 3 and r[0] == 'Line No':
        hdr = r; ii = hdr.index('Instructions Executed'); ti = hdr.index('Thread Instructions Executed'); si = hdr.index('# Samples'); continue
    if hdr is None or len(r) <= ii or r[0] == '': continue
    key = (cur, num(r[0])); agg[key] += num(r[ii]); src[key] = r[1].strip()[:120]; samp[key] += num(r[si]); thr[key] += num(r[ti])
tot = sum(agg.values()); ts = sum(samp.values())
print("total warp-instructions", tot, "stall samples", ts)
byfile = collections.Counter()
for k, v in agg.items(): byfile[k[0]] += v
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.most_common()})
for k, v in agg.most_common(top):
    print("%-12s %4d %5.1f%% smp %4.1f%% act %4.1f | %s" % (k[0], k[1], 100 * v / tot, 100 * samp[k] / max(ts, 1), thr[k] / max(v, 1), src[k]))
