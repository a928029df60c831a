"""Summarise an ncu --set full capture: key raw metrics + per-SASS-region instruction/stall shares.
usage: python tools/ncu_summary.py raw.csv src.csv"""
import csv
import sys
from collections import Counter

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = [
    "gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
isrc, isamp, iinst = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]
tot_inst = sum(int(r[iinst] or 0) for r in data)
tot_s = sum(int(r[isamp] or 0) for r in data)
print("total warp-instructions", tot_inst, "samples", tot_s)
regions, cur = [], None
for k, r in enumerate(data):
    inst, s = int(r[iinst] or 0), int(r[isamp] or 0)
    lvl = inst / tot_inst * 100
    if cur is None or abs(lvl - cur[2]) > 0.05 * max(lvl, cur[2], 0.02):
        cur = [k, k, lvl, 0, 0, []]
        regions.append(cur)
    cur[1] = k
    cur[3] += inst
    cur[4] += s
    tok = r[isrc].split()
    cur[5].append(tok[1] if tok and tok[0].startswith("@") else (tok[0] if tok else ""))
for a, b, lvl, inst, s, ops in regions:
    if inst / tot_inst < 0.004 and s / tot_s < 0.004:
        continue
    print(f"[{a:3d}-{b:3d}] n={b-a+1:3d} each {lvl:5.2f}%  inst {inst/tot_inst*100:5.1f}%  stall-samples {s/tot_s*100:5.1f}%  {Counter(ops).most_common(7)}")
