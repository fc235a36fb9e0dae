"""Summarise an ncu --set full report (exported with `ncu -i X.ncu-rep --page raw --csv`) per launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]
cols = []
for w in want:
    m = [h for h in hdr if h.endswith(w) or h == w]
    if m:
        cols.append(m[0])
print("launch | kernel | grid | " + " | ".join(c.split(".TriageCompute.")[-1] for c in cols))
for n, d in enumerate(data):
    print(n, "|", d[idx["Kernel Name"]].split("(")[0][-40:], "|", d[idx["Grid Size"]], "|",
          " | ".join(f"{d[idx[c]]} {units[idx[c]]}" for c in cols))
