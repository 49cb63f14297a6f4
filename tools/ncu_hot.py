"""Per-region stall breakdown of one kernel from an ncu report (source page, SASS): which warp role waits on what.
Usage: python tools/ncu_hot.py <report.ncu-rep> [first_instr last_instr]"""
import csv, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
cols = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    data.append(r)
lo, hi_ = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, len(data))
tot = {h: 0 for h in stall}
for r in data[lo:hi_]:
    for h in stall:
        tot[h] += int(r[cols[h]] or 0)
n = sum(tot.values())
print("instructions %d..%d of %d, %d samples" % (lo, hi_, len(data), n))
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print("  %-24s %7d  %5.1f%%" % (h, v, 100.0 * v / max(n, 1)))
top = sorted(range(lo, hi_), key=lambda i: -int(data[i][cols["# Samples"]] or 0))[:12]
for i in top:
    r = data[i]
    print("  [%5d] %-70s samples %6s exec %s" % (i, r[cols["Source"]].strip()[:70], r[cols["# Samples"]], r[cols["Instructions Executed"]]))
