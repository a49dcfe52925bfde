"""Top-sampled SASS lines of one kernel in an .ncu-rep: python tools/ncu_hot.py rep kernel_regex [launch_idx] [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
idx = sys.argv[3] if len(sys.argv) > 3 else "1"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", "::regex:%s:%s" % (rx, idx)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    if r[hdr.index('# Samples')] == '# Samples':
        break
    data.append(r)
si = hdr.index("# Samples"); so = hdr.index("Source"); ie = hdr.index("Instructions Executed")
tot = sum(int(r[si] or 0) for r in data)
print(rows[0][1][:100], "total samples", tot, "SASS lines", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:top]
for i in sorted(order):
    r = data[i]
    print("%5d %6s %5.1f%% exec=%-8s %s" % (i, r[si], 100.0 * int(r[si] or 0) / max(tot, 1), r[ie], r[so].strip()[:110]))
