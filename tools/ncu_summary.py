"""Summarise an `ncu --page raw --csv` export into the metric,unit,value table committed under profiles/.
usage: python tools/ncu_summary.py raw.csv "header comment" > profiles/ncu_<kernel>_<round>_summary.csv"""
import csv
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
for vals in rows[2:]:
    if "Kernel Name" in hdr:
        print("# kernel: " + vals[hdr.index("Kernel Name")])
    print("metric,unit,value")
    for i, h in enumerate(hdr):
        if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            print("%s,%s,%s" % (h, units[i], vals[i]))
