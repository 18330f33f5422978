"""Markdown summary of an ncu report: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/xxx.md

One table per captured launch with the metrics the roofline discussion uses (read with `ncu -i ... --page raw --csv`), plus the
top warp-stall reasons of the first launch of each report."""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        print(f"## {rep.split('/')[-1]}\n")
        for n, r in enumerate(data, 1):
            print(f"#### launch {n}: `{r[col['Kernel Name']]}`  grid {r[col['Grid Size']]} x block {r[col['Block Size']]}\n")
            print("| metric | unit | value |\n|---|---|---|")
            for m in METRICS:
                if m in col:
                    print(f"| {m} | {units[col[m]]} | {r[col[m]]} |")
            print()
        stalls = {h: i for h, i in col.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
        if stalls and data:
            vals = []
            for h, i in stalls.items():
                try:
                    vals.append((float(data[0][i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
            tot = sum(v for v, _ in vals) or 1.0
            top = sorted(vals, reverse=True)[:9]
            print("Warp-stall samples, launch 1: " + ", ".join(f"{name} {100 * v / tot:.0f}%" for v, name in top) + "\n")


if __name__ == "__main__":
    main()
