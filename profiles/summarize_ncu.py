"""Turn an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.

usage: python profiles/summarize_ncu.py gpurun_out/X.ncu-rep > profiles/r01_X.txt
Reads the report with `ncu -i ... --page raw --csv` (and the source page for the hottest SASS lines)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}  (ncu --set full --clock-control none; profiler timings are cold-cache, serialised)")
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print(f"\n## kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"{k:70s} {d[k]:>18s} {u.get(k, '')}")
        stall = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                        if k.startswith("smsp__average_warp") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)[:6]
        if not stall:
            stall = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                            if "warp_issue_stalled" in k and k.endswith(".ratio") and v not in ("", "n/a")), reverse=True)[:6]
        for v, k in stall:
            print(f"{k:70s} {v:18.3f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    try:
        h = next(i for i, r in enumerate(rows) if "# Samples" in r)
    except StopIteration:
        return
    hdr = rows[h]
    isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = [r for r in rows[h + 1:] if len(r) > isamp and r[isamp].isdigit()]
    tot = sum(int(r[isamp]) for r in data) or 1
    print(f"\n## hottest SASS lines (of {tot} samples, last kernel in the report)")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:16]:
        print(f"{100 * int(r[isamp]) / tot:5.1f}%  exec {r[iex]:>12s}  {r[isrc].strip()[:110]}")


if __name__ == "__main__":
    main()
