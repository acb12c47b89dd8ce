#!/bin/bash
# Round 2, call p (N GPUs): sharded parity (incl. the periodic block path) and the bench lines at L = 32 (with e2e through
# the copy engine) and L = 34.  Usage: gpurun --gpus N -- 'bash scripts/gpu_r2p.sh <tag> N "32 34"'
TAG=${1:-r2p}; N=${2:-2}; LS=${3:-"32 34"}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print("ms/apply", round(d["ms_per_step"],3), "parity", d.get("parity"), "e2e ms", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception: pass'
timeout 300 $TR scripts/mgpu_check.py > $O/mgpu_${TAG}_n${N}.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}_n${N}.log
echo "== mgpu_check N=$N: $(grep -h 'FAIL\|ALL OK\|rc=' $O/mgpu_${TAG}_n${N}.log | tr '\n' ' ')" | tee -a $O/r2p_${TAG}_n${N}.txt
nvidia-smi nvlink -gt d -i 0 > $O/nvlink_before_${TAG}_n${N}.txt 2>&1
for L in $LS; do
  E2E=""; if [ "$L" != "32" ]; then E2E="--no-e2e"; fi
  timeout 400 $TR bench.py --gpus $N --L $L --steps 10 --warmup 3 --no-cpu --no-solve $E2E > $O/bench_${TAG}_n${N}_L$L.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}_n${N}_L$L.log
  echo "N=$N L=$L: $(tail -n 2 $O/bench_${TAG}_n${N}_L$L.log | python -c "$J")" | tee -a $O/r2p_${TAG}_n${N}.txt
  grep "^{" $O/bench_${TAG}_n${N}_L$L.log | tail -n 1 > $O/benchline_${TAG}_n${N}_L$L.json
done
nvidia-smi nvlink -gt d -i 0 > $O/nvlink_after_${TAG}_n${N}.txt 2>&1
python - <<PY | tee -a $O/r2p_${TAG}_n${N}.txt
import re
def tot(f):
    rx = tx = 0
    for l in open(f):
        m = re.search(r"Data (Rx|Tx): (\d+) KiB", l)
        if m:
            if m.group(1) == "Rx": rx += int(m.group(2))
            else: tx += int(m.group(2))
    return rx, tx
try:
    b, a = tot("$O/nvlink_before_${TAG}_n${N}.txt"), tot("$O/nvlink_after_${TAG}_n${N}.txt")
    print("NVLink GPU0 over all bench runs above: rx %.2f GB tx %.2f GB" % ((a[0] - b[0]) * 1024 / 1e9, (a[1] - b[1]) * 1024 / 1e9))
except Exception as e:
    print("nvlink counters unavailable:", e)
PY
