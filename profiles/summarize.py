"""Summaries kept under profiles/ from ncu artefacts (gpurun_out/ is scratch).

  python profiles/summarize.py launches  gpurun_out/launches_r01b.csv           > profiles/r01b_launch_share.txt
  python profiles/summarize.py full      gpurun_out/prof_r1i.ncu-rep [...]      > profiles/r01b_ncu_full.csv
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = r[kn].split("(")[0].replace("void ", "").replace("unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none: per-kernel launches, mean duration, share of the captured time")
    print("# (serialised, cold-cache replays: the SHARE is what must agree with the CUDA-event times of bench.py)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s launches %4d  mean %9.1f us  share %5.1f %%" % (k[:60], n, t / n / 1e3, 100 * t / tot))


def full(paths):
    w = csv.writer(sys.stdout)
    w.writerow(["report", "kernel"] + KEEP)
    for p in paths:
        txt = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr = rows[0]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            w.writerow([p.split("/")[-1], d["Kernel Name"][:70]] + [d.get(k, "") for k in KEEP])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])
