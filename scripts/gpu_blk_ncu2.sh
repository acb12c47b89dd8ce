#!/bin/bash
# time + dram bytes of the block kernel under env settings. Usage: gpu_blk_ncu2.sh <tag> "ENV..." ...
TAG=$1; shift; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve"
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-solve"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed.avg.per_cycle_elapsed,smsp__inst_executed.sum
for e in "$@"; do
  r=$(env $e timeout 200 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])" 2>&1 | tail -1)
  env $e timeout 600 ncu --metrics $M --clock-control none -k regex:sd_blkl?_apply -s 1 -c 1 --csv --log-file $O/n2.csv $P > /dev/null 2>&1
  d=$(python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/n2.csv")) if len(r)>10]
v={r[-3]:float(r[-1].replace(',','')) for r in rows[1:]}
print(f"dramR={v['dram__bytes_read.sum']/1e9:.1f}GB W={v['dram__bytes_write.sum']/1e9:.1f}GB hit={v['lts__t_sector_hit_rate.pct']:.0f}% lsu={v['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']:.0f}% ipc={v['sm__inst_executed.avg.per_cycle_elapsed']:.2f} inst={v['smsp__inst_executed.sum']/1e9:.2f}e9")
PY
)
  echo "$e -> $r ms $d" | tee -a $O/blkenv_${TAG}.txt
done
