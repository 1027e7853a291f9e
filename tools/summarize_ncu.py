"""Turns ncu outputs (launch-list CSV, .ncu-rep) into the small text summaries kept under profiles/."""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    out = [f"# per-kernel device time of ONE eager forward->mask step (ncu gpu__time_duration.sum, cold cache, serialised)",
           f"# source: {path}; compare SHARES, not absolutes", f"{'kernel':62s} {'n':>4s} {'us':>10s} {'share':>7s}"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{k:62s} {n:4d} {us:10.1f} {100 * us / total:6.1f}%")
    out.append(f"{'TOTAL':62s} {sum(v[0] for v in agg.values()):4d} {total:10.1f}")
    return "\n".join(out)


def report(path, max_kernels=3):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
    out = [f"# ncu --set full summary of {path}"]
    for r in rows[2:2 + max_kernels]:
        out.append("kernel: " + r[idx["Kernel Name"]][:110])
        for k in KEYS:
            if k in idx:
                out.append(f"  {k} [{units[idx[k]]}] = {r[idx[k]]}")
        st = sorted(((float(r[idx[h]]), h) for h in stall if r[idx[h]] not in ("", "n/a")), reverse=True)[:5]
        out.append("  top stalls (warps per issue): " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}" for v, h in st))
    return "\n".join(out)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    print(launches(path) if mode == "launches" else report(path))
