"""Groups the SASS lines of an ncu source page by execution count (= loop nest) and prints, per group,
its share of executed instructions, of stall samples, and its stall mix."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
groups = collections.OrderedDict()
for k, r in enumerate(data):
    ex = int(r[iex])
    key = None
    for g in groups:
        if abs(ex - g) <= 0.03 * max(g, 1):
            key = g
            break
    if key is None:
        key = ex
        groups[key] = [0, 0, collections.Counter(), k, k, 0]
    g = groups[key]
    g[0] += ex
    g[1] += int(r[ismp])
    g[4] = k
    g[5] += 1
    for i, h in stall:
        g[2][h] += int(r[i])
ti = sum(g[0] for g in groups.values())
ts = sum(g[1] for g in groups.values())
print("total warp instructions %.3e, samples %d" % (ti, ts))
for key, g in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    if g[1] < 0.005 * ts:
        continue
    mix = ", ".join("%s %.0f%%" % (h[6:], 100.0 * c / max(1, sum(g[2].values()))) for h, c in g[2].most_common(5))
    print("exec %9.1fM x %4d instr (sass %5d..%5d): inst %5.1f%%  samples %5.1f%%  | %s" % (
        key / 1e6, g[5], g[3], g[4], 100.0 * g[0] / ti, 100.0 * g[1] / ts, mix))
