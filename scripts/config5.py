"""BASELINE.json config 5: XXZChain L=36 nup=18 (9.08e9 states, u64 ranks), chebyshev_time_evolve cheb_n=100 dt=0.1 with
given Ebounds, c128, sharded over the GPUs of one box (Chebyshev.jl:61-124 through sd_chebyshev_evolve).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        scripts/config5.py [--L 36] [--cheb-n 100] [--dt 0.1]

Prints one JSON line: wall time of the evolution, ms per Chebyshev term, norm conservation | ||psi(t)|| / ||psi0|| - 1 |,
and (small L only, --check) the relative L2 distance to the oracle's psi(t)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import spindyn as sd  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=36)
ap.add_argument("--cheb-n", type=int, default=100)
ap.add_argument("--dt", type=float, default=0.1)
ap.add_argument("--check", action="store_true", help="compare with the oracle (L <= 24)")
args = ap.parse_args()

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = sd.Context.from_torch_distributed(local_rank)
else:
    ctx = sd.Context(0)
rank = ctx.rank
sd.set_default_context(ctx)
L, nup = args.L, args.L // 2
t0 = time.perf_counter()
model = sd.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=nup, ctx=ctx)
N = model.dim
first, count = model.local_range
seed = 20261018
psi0 = model.vector(np.complex128).fill_seeded(seed, 1.0 / np.sqrt(2.0 * N / 3.0))
n0 = psi0.norm()
t_setup = time.perf_counter() - t0
# open Heisenberg chain: Emax = (L - 1) / 4 exactly, E0 > -0.4432 L; generous given bounds (the reference takes them as input)
Ebounds = (-0.4432 * L - 0.5, (L - 1) / 4 + 0.25)
ctx.sync()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
l0 = ctx.launch_count()
psit = sd.chebyshev_time_evolve(psi0, args.dt, sd.apply_H_, model, cheb_n=args.cheb_n, Ebounds=Ebounds, device=True)
ctx.sync()
wall = time.perf_counter() - t0
n1 = psit.norm()
res = {"config": f"XXZChain L={L} nup={nup} c128 chebyshev_time_evolve cheb_n={args.cheb_n} dt={args.dt} Ebounds={Ebounds}",
       "n_gpus": world, "states": int(N), "bytes_per_vector": int(N) * 16, "rank_bits": model.info.get("rank_bits"),
       "kernel_path": model.info["kernel_path"], "setup_s": t_setup, "wall_s": wall, "ms_per_term": wall * 1e3 / args.cheb_n,
       "launches_rank0": ctx.launch_count() - l0, "norm0": n0, "norm_t": n1, "norm_drift": abs(n1 / n0 - 1.0)}
if args.check:
    from oracle import oracle as orc
    om = orc.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=nup)
    ref0 = orc.fill_seeded(N, seed, cplx=True) * (1.0 / np.sqrt(2.0 * N / 3.0))
    ref = orc.chebyshev_time_evolve(ref0, args.dt, orc.apply_H_, om, cheb_n=args.cheb_n, Ebounds=Ebounds)
    got = psit.to_host()
    err2 = torch.tensor([float(np.sum(np.abs(got - ref[first:first + count]) ** 2))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(err2)
    res["rel_l2_vs_oracle"] = float(np.sqrt(err2.item()) / np.linalg.norm(ref))
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
