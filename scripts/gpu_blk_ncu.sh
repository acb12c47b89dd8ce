#!/bin/bash
# a few ncu counters of the block kernel under several SD_BLK_DBG settings. Usage: gpu_blk_ncu.sh <tag> [dbg list]
TAG=${1:-n}; shift; O=gpurun_out; mkdir -p $O
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_read.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct
for dbg in ${@:-0}; do
  SD_BLK_DBG=$dbg timeout 600 ncu --metrics $M --clock-control none -k regex:sd_blkr?_apply -s 1 -c 1 --csv --log-file $O/ncum_${TAG}_$dbg.csv $P > /dev/null 2>&1
  echo "== dbg=$dbg"; python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/ncum_${TAG}_$dbg.csv")) if len(r)>10]
for r in rows[1:]: print(f"{r[-3]:75s} {r[-1]:>18s} {r[-2]}")
PY
done
