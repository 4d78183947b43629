"""Summarise an .ncu-rep (read with `ncu -i`) into the handful of metrics DESIGN.md / bench.py cite.

    python profiles/summarize.py gpurun_out/decode_attn_r1.ncu-rep > profiles/r1_decode_attn.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__cluster_size", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    print(f"# ncu summary of `{path}`\n")
    for n, r in enumerate(rows[2:]):
        print(f"## launch {n}: `{r[name_i][:100]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for i, h in enumerate(hdr):
            if h in KEYS:
                print(f"| {h} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
