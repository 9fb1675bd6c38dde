"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.
usage: python profiles/summarize.py launches <launches.csv> | full <report.ncu-rep>"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "lts__t_bytes.sum", "lts__t_sectors.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9, "s": 1e9}.get(r[ui], 1.0)
        agg[r[ki][:70]][0] += 1
        agg[r[ki][:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"launches {sum(v[0] for v in agg.values())}  total {tot / 1e6:.3f} ms (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{v[1] / 1e6:10.3f} ms {v[0]:6d} x {100 * v[1] / tot:5.1f}%  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name", "?")[:90])
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                print(f"  {h} = {v} {u}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
