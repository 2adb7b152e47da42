"""Turn ncu outputs into the small text summaries committed under profiles/.
  python tools/ncu_summary.py launches <launches.csv>            -> per-kernel launch counts / total device time
  python tools/ncu_summary.py kernel <file.ncu-rep> [kernel idx] -> key counters of one captured kernel
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("atlas::<unnamed>::", "").replace("void ", "")[-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * (1e-3 if r[ui] == "ns" else 1.0)
    total = sum(t for _, t in agg.values())
    print(f"{'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>6s}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:8d} {t:10.1f} {t / n:9.1f} {100 * t / total:5.1f}%  {k}")
    print(f"{'':8s} {total:10.1f} us total (ncu-serialised, cold cache: compare shares, not absolutes)")


def kernel(path, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    V = rows[2 + which]
    print("kernel:", V[H.index("Kernel Name")][:100])
    for k in KEYS:
        if k in H:
            i = H.index(k)
            print(f"  {k:75s} {V[i]:>16s} {U[i]}")
    for i, h in enumerate(H):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(V[i])
            except ValueError:
                continue
            if v >= 0.3:
                print(f"  stall {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):40s} {v:6.2f} warps per issue-active cycle")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        kernel(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
