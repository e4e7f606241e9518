"""Turn an ncu report (gpurun_out/*.ncu-rep, scratch) into the small text summaries committed under profiles/.

    python profiles/summarize.py gpurun_out/fill_v3.ncu-rep profiles/r01_fill_v3   [items_in_capture]

Writes <out>_metrics.txt (selected `--page raw` metrics) and <out>_hotspots.txt (instructions executed and stall
samples per 256-byte SASS window, from `--page source`; compile with -lineinfo)."""
import collections
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.avg.per_second", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    items = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    with open(out + "_metrics.txt", "w") as f:
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"# kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\n")
            for h, u, v in zip(hdr, units, r):
                if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                    f.write(f"{h:95s} {v:>18s} {u}\n")
    rows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    name = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr, data = rows[1], rows[2:]
    ia, ie, ism, ith = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    base = int(data[0][ia], 16)
    tot = sum(int(r[ie]) for r in data)
    win = collections.OrderedDict()
    for r in data:
        k = (int(r[ia], 16) - base) // 0x100
        w = win.setdefault(k, [0, 0, 0])
        w[0] += int(r[ie]); w[1] += int(r[ism]); w[2] += int(r[ith])
    with open(out + "_hotspots.txt", "w") as f:
        f.write(f"# {name}\n# total warp instructions executed: {tot}")
        if items:
            f.write(f"  ({tot / items:.0f} per item, {items:g} items in this capture)")
        f.write("\n# window  warp_instr  share  stall_samples  active_lanes/32\n")
        for k, (i, sm, th) in win.items():
            if i:
                f.write(f"{k * 0x100:#7x} {i:12d} {100 * i / tot:5.1f}% {sm:8d} {th / i / 32:5.2f}\n")


if __name__ == "__main__":
    main()
