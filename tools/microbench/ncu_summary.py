"""Summarises `ncu -i X.ncu-rep --page raw --csv` (one row per captured launch) into the few metrics DESIGN.md quotes.
Usage: python tools/microbench/ncu_summary.py raw.csv > profiles/NAME.txt"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum", "lts__t_sector_hit_rate.pct"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("kernel:", r[ix["Kernel Name"]][:90])
    for w in WANT:
        if w in ix:
            print("  %-70s %s %s" % (w, r[ix[w]], units[ix[w]]))
    for h in hdr:  # DMMA runs on the tensor pipe: its activity is not in sm__pipe_fp64_*
        if ("pipe_tensor_cycles_active_realtime.avg.pct" in h or "subpipe_dmma_cycles_active.avg" in h) and r[ix[h]]:
            print("  %-70s %s %s" % (h.split("TriageCompute.")[-1], r[ix[h]], units[ix[h]]))
    st = {h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""): float(r[ix[h]]) for h in hdr
          if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and r[ix[h]]}
    print("  stalled warps per issue: " + " ".join("%s=%.2f" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
