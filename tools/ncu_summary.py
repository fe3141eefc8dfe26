"""Print the roofline-relevant raw metrics of every kernel in an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "sm__sass_inst_executed_op_shared.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print("==", r[ix["Kernel Name"]][:90], "grid", r[ix["Grid Size"]], "block", r[ix["Block Size"]])
    for k in KEYS:
        if k in ix:
            print(f"   {k:75s} {r[ix[k]]:>16s} {units[ix[k]]}")
