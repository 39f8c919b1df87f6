"""Rank CUDA source lines of an ncu report by executed warp instructions / stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == 'Line No')
start = rows.index(hdr) + 1
res = []
for r in rows[start:]:
    if r and r[0] not in ('', 'Line No'):
        try: res.append((int(r[7]), int(r[6]), r[0], r[1].strip()[:120]))
        except (ValueError, IndexError): pass
tot = sum(o[0] for o in res); ts = sum(o[1] for o in res)
print("total warp instructions", tot, "samples", ts)
key = 1 if (len(sys.argv) > 3 and sys.argv[3] == 'samples') else 0
res.sort(key=lambda o: -o[key])
for o in res[:top]:
    print("%11d %5.1f%% samp %6d %5.1f%%  L%-4s %s" % (o[0], 100 * o[0] / tot, o[1], 100 * o[1] / max(ts, 1), o[2], o[3]))
