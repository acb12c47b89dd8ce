#!/bin/bash
# A/B of the halo mirror (SD_HALO=1: chunked copy-engine prefetch of the peer ranges into a sparse local mapping,
# sd_halo_host.h) on N GPUs of one box: sharded parity under both settings, then the L=32 bench.  Every step has its own
# short timeout.  Usage: gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_halo_ab.sh <tag> 2'   (then 4, then 8)
TAG=$1; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for halo in 0 1; do
  for extra in "" "SD_BLK_RING=1"; do
    tag=${TAG}_n${N}_h${halo}_${extra:+ring}
    env SD_HALO=$halo $extra timeout 240 $TR scripts/mgpu_check.py 32 > $O/mgpu_${tag}.log 2>&1; echo "rc=$?" >> $O/mgpu_${tag}.log
    env SD_HALO=$halo $extra timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $O/bench_${tag}.log 2>&1; echo "rc=$?" >> $O/bench_${tag}.log
    echo "== N=$N SD_HALO=$halo $extra: $(grep -h 'FAIL\|ALL OK\|rc=' $O/mgpu_${tag}.log | tail -n 2 | tr '\n' ' ') $(tail -n 2 $O/bench_${tag}.log | python -c "import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('ms/apply', round(d['ms_per_step'],3))
    except Exception: pass")" | tee -a $O/halo_${TAG}.txt
  done
done
# shards weighted by remote volume (sd_halo_balance): meant for 8 ranks, where the 101 / 010 ranks pull 2.5 shards
for cost in 70 100 140; do
  tag=${TAG}_n${N}_bal${cost}
  env SD_HALO=1 SD_SHARD_BALANCE=1 SD_SHARD_REMOTE_COST=$cost timeout 240 $TR scripts/mgpu_check.py > $O/mgpu_${tag}.log 2>&1; echo "rc=$?" >> $O/mgpu_${tag}.log
  r=$(env SD_HALO=1 SD_SHARD_BALANCE=1 SD_SHARD_REMOTE_COST=$cost timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])" 2>&1 | tail -n 1)
  echo "N=$N SD_HALO=1 SD_SHARD_BALANCE=1 remote_cost=$cost -> $r ms; parity: $(grep -h 'FAIL\|ALL OK\|rc=' $O/mgpu_${tag}.log | tail -n 2 | tr '\n' ' ')" | tee -a $O/halo_${TAG}.txt
done
for chunks in 4 16 32; do
  r=$(env SD_HALO=1 SD_HALO_CHUNKS=$chunks timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])" 2>&1 | tail -n 1)
  echo "N=$N SD_HALO=1 chunks=$chunks -> $r ms" | tee -a $O/halo_${TAG}.txt
done
