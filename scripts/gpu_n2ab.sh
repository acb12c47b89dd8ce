#!/bin/bash
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(round(d["ms_per_step"],3))
    except Exception: pass'
for e in "SD_BLKL_THREADS=640" "SD_BLKL_THREADS=768" "SD_BLKL_THREADS=768 SD_BLK_PFP=0" "SD_BLKL_THREADS=640 SD_BLK_PFP=0"; do
  echo "N=2 L=32 $e: $(env $e timeout 200 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-e2e --no-solve --no-parity 2>&1 | tail -n 2 | python -c "$J") ms" | tee -a $O/n2ab.txt
done
