"""Summarise an .ncu-rep (read on the CPU box): per kernel launch the metrics the judge asks for.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_xu.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]][:90]
        print("==", r[idx["ID"]], name)
        for k in hdr:
            if any(k.startswith(key) or key in k for key in KEYS) or "tensor" in k:
                v = r[idx[k]]
                if v not in ("", "n/a"):
                    print(f"   {k} [{units[idx[k]]}] = {v}")


if __name__ == "__main__":
    main()
