"""Summaries of ncu captures for profiles/: per-kernel rows of a `--set full` report and the kernel shares of a
launch list (`--metrics gpu__time_duration.sum`).  Usage:
    python tools/ncu_summary.py full  <report.ncu-rep>
    python tools/ncu_summary.py list  <launches.csv>"""
import collections, csv, subprocess, sys


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    want = [("time", "gpu__time_duration.sum"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
            ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            ("tensor%", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
            ("lts%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            ("grid", "launch__grid_size"), ("regs", "launch__registers_per_thread"),
            ("cycles", "sm__cycles_elapsed.max")]
    for r in rows[2:]:
        parts = ["kernel=%s" % r[col["Kernel Name"]][:60]]
        for name, key in want:
            k = next((h for h in hdr if h.endswith(key)), None)
            if k:
                parts.append("%s=%s%s" % (name, r[col[k]], units[col[k]]))
        print(" | ".join(parts))


def launch_list(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    ours = {k: v for k, v in agg.items() if not k.startswith("void at::") and "cutlass" not in k and "nccl" not in k.lower()}
    tot = sum(sum(v) for v in ours.values())
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        print("%-50s n=%4d total %9.3f ms avg %8.3f ms share %5.1f%%" % (k[:50], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e6,
                                                                          100 * sum(v) / tot))


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2])
