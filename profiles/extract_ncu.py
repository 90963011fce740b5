#!/usr/bin/env python
"""Summarise an `ncu --set full` report (run here, no GPU needed):
    python profiles/extract_ncu.py gpurun_out/r1_prof_tf32x3.ncu-rep > profiles/r1_ncu_tf32x3_summary.txt
and a launch list (`ncu --metrics gpu__time_duration.sum --csv`):
    python profiles/extract_ncu.py --launches gpurun_out/r1_launches_fused.csv > profiles/r1_launches_fused_summary.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print("  %-88s %-10s %s" % (k, units[i], r[i]))


def launches(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) < 15:
            continue
        name = r[4].split("(")[0][-60:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(v[1] for v in agg.values())
    print("%10s %6s %6s  kernel   (gpu__time_duration.sum per launch, cold cache, serialised: compare SHARES)" % ("total ms", "count", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.3f %6d %5.1f%%  %s" % (v[1] / 1e6, v[0], 100 * v[1] / tot, k))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1])
