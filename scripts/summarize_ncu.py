"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python scripts/summarize_ncu.py full gpurun_out/prof_p2.ncu-rep profiles/r01_full_2d_p2.md
"""
import collections
import csv
import io
import subprocess
import sys

FULL_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum",
    "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_fp64.sum",
    "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_generic_ld.sum",
    "smsp__inst_executed_op_global_ld.sum",
]


def short(name):
    return name.replace("void ", "").split("(")[0]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        k = short(r[ik])
        t = tot.setdefault(k, [0, 0.0, r[ig], r[ib]])
        t[0] += 1
        t[1] += float(r[iv]) / 1e3
    total = sum(t[1] for t in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` - per-launch device time, cold cache, serialised; "
                "compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | avg us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, (n, us, g, b) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {us:.1f} | {us / n:.2f} | {100 * us / total:.1f}% | {g} | {b} |\n")
        f.write(f"\ntotal {total:.1f} us over {sum(t[0] for t in tot.values())} launches\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\nOne column per captured launch.\n\n")
        names = [short(r[ik]) for r in rows[2:]]
        f.write("| metric | unit | " + " | ".join(f"`{n}`" for n in names) + " |\n")
        f.write("|---|---|" + "---:|" * len(names) + "\n")
        for k in FULL_KEYS:
            if k not in hdr:
                continue
            i = hdr.index(k)
            f.write(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
