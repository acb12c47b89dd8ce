#!/bin/bash
# Round 2, call d (2 GPUs): straight-order lean kernel with L1 prefetch of the remote entries: 1-GPU check, sharded
# parity + timing, bench line with parity at N=2, config 5 in small with the oracle check, L=34 on 1 and 2 GPUs.
TAG=${1:-r2d}; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print("ms/apply", round(d["ms_per_step"],3), "parity", d.get("parity"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception: pass'
echo "1 GPU L=32: $(timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -1 | python -c "$J")" | tee -a $O/r2d_${TAG}.txt
echo "1 GPU L=34: $(timeout 300 python bench.py --L 34 --steps 5 --warmup 2 --no-cpu --no-e2e --no-solve 2>&1 | tail -1 | python -c "$J")" | tee -a $O/r2d_${TAG}.txt
timeout 300 $TR scripts/mgpu_check.py 32 > $O/mgpu_${TAG}_n${N}.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}_n${N}.log
echo "== N=$N: $(grep -h 'FAIL\|ALL OK\|ms/apply\|rc=' $O/mgpu_${TAG}_n${N}.log | tr '\n' ' ')" | tee -a $O/r2d_${TAG}.txt
for L in 32 34; do
  timeout 300 $TR bench.py --gpus $N --L $L --steps 10 --warmup 3 --no-cpu --no-solve > $O/bench_${TAG}_n${N}_L$L.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}_n${N}_L$L.log
  echo "N=$N L=$L: $(tail -n 2 $O/bench_${TAG}_n${N}_L$L.log | python -c "$J")" | tee -a $O/r2d_${TAG}.txt
done
timeout 300 $TR scripts/config5.py --L 24 --check > $O/config5_${TAG}_n${N}.log 2>&1; echo "rc=$?" >> $O/config5_${TAG}_n${N}.log
tail -n 2 $O/config5_${TAG}_n${N}.log | cut -c1-900 | tee -a $O/r2d_${TAG}.txt
