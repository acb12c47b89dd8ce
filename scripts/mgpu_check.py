"""Sharded (one rank per GPU) parity check against the CPU oracle, run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 scripts/mgpu_check.py [L_time ...]

Every rank builds the same seeded psi, uploads its own rank range, runs the sharded
H.psi / recurrences (peer shards are read over NVLink) and compares its slice of the
result with the oracle's full-vector answer.  Optional arguments: chain lengths to time.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import numpy as np
import torch
import torch.distributed as dist
import spindyn as sd
from oracle import oracle as orc

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = sd.Context.from_torch_distributed(local_rank)
sd.set_default_context(ctx)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def allmax(x):
    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


ok = True


def report(name, err, tol):
    global ok
    e = allmax(err)
    good = e < tol
    ok &= good
    if rank == 0:
        print(f"[world={world}] {name}: err={e:.2e} tol={tol:.0e} {'OK' if good else 'FAIL'}", flush=True)


cases = [(16, 8, "open"), (18, 9, "open"), (20, 10, "open"), (20, 7, "open"), (17, 8, "open"), (16, 8, "periodic"), (18, 7, "periodic"), (12, None, "open")]
for (L, nup, bc) in cases:
    for dtype in (np.float64, np.complex128):
        kw = dict(Jxy=0.7, Jz=1.3, hz=0.2, nup=nup, boundary=bc)
        m = sd.XXZChain(L, ctx=ctx, **kw)
        om = orc.XXZChain(L, **kw)
        N = m.dim
        first, count = m.local_range
        rng = np.random.default_rng(L * 100 + (nup or 0))
        psi = rng.standard_normal(N).astype(dtype)
        if dtype == np.complex128:
            psi = psi + 1j * rng.standard_normal(N)
        ref = np.empty_like(psi)
        orc.apply_H_(ref, psi, om)
        d = m.to_device(psi[first:first + count])
        o = m.vector(dtype)
        sd.apply_H_(o, d, m)
        got = o.to_host()
        err = np.linalg.norm(got - ref[first:first + count]) / np.linalg.norm(ref)
        report(f"apply_H L={L} nup={nup} {bc} {np.dtype(dtype).name} path={m.info['kernel_path']} local={count}", err, 1e-13)
        # round trip + global reductions
        back = d.to_host()
        report("  upload/download round trip", 0.0 if np.array_equal(back, psi[first:first + count]) else 1.0, 0.5)
        report("  norm (NCCL all-reduce)", abs(d.norm() - np.linalg.norm(psi)) / np.linalg.norm(psi), 1e-13)
        report("  dot", abs(d.dot(o) - np.vdot(psi, ref)) / abs(np.vdot(psi, ref)), 1e-12)
        del d, o, m

# ---- recurrences on a sharded vector (config 0's model)
L, nup = 16, 8
m = sd.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=nup, ctx=ctx)
om = orc.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=nup)
N = m.dim
first, count = m.local_range
rng = np.random.default_rng(7)
v0 = rng.standard_normal(N)
E_ref, psi_ref = orc.lanczos_groundstate(orc.apply_H_, om, lanc_m=100, v0=v0.copy())
E, psi = sd.lanczos_groundstate(sd.apply_H_, m, lanc_m=100, v0=m.to_device(v0[first:first + count]), device=True)
report(f"lanczos_groundstate E0 (ref {E_ref:.12f})", abs(E - E_ref), 1e-10)
g = psi.to_host()
sgn = 1.0 if allmax(0.0) == 0.0 else 1.0
# fix the global sign with rank 0's first element of the reference overlap
ov = torch.tensor([float(np.dot(g, psi_ref[first:first + count]))], device="cuda", dtype=torch.float64)
dist.all_reduce(ov)
report("lanczos_groundstate |<psi|psi_ref>| - 1", abs(abs(float(ov.item())) - 1.0), 1e-9)

phi = rng.standard_normal(N) + 1j * rng.standard_normal(N)
phi /= np.linalg.norm(phi)
Emin, Emax = -8.0, 5.0
a, b = (Emax - Emin) / 2 * 1.01, (Emax + Emin) / 2
mu_ref = orc.compute_chebyshev_moments(orc.apply_H_, phi.copy(), 64, a, b, om)
mu = sd.compute_chebyshev_moments(sd.apply_H_, m.to_device(phi[first:first + count]), 64, a, b, m)
report("KPM moments (64)", np.max(np.abs(np.asarray(mu) - np.asarray(mu_ref))), 1e-11)

psi0 = np.zeros(N, dtype=np.complex128)
psi0[:] = rng.standard_normal(N) + 1j * rng.standard_normal(N)
psi0 /= np.linalg.norm(psi0)
pt_ref = orc.chebyshev_time_evolve(psi0.copy(), 0.3, orc.apply_H_, om, cheb_n=40, Ebounds=(Emin, Emax))
pt = sd.chebyshev_time_evolve(m.to_device(psi0[first:first + count]), 0.3, sd.apply_H_, m, cheb_n=40, Ebounds=(Emin, Emax), device=True)
report("chebyshev_time_evolve", np.linalg.norm(pt.to_host() - pt_ref[first:first + count]) / np.linalg.norm(pt_ref) * np.sqrt(world), 1e-9)
kt_ref = orc.krylov_time_evolve(psi0.copy(), 0.3, orc.apply_H_, om, kry_m=20)
kt = sd.krylov_time_evolve(m.to_device(psi0[first:first + count]), 0.3, sd.apply_H_, m, kry_m=20, device=True)
report("krylov_time_evolve", np.linalg.norm(kt.to_host() - kt_ref[first:first + count]) / np.linalg.norm(kt_ref) * np.sqrt(world), 1e-9)

if rank == 0:
    print("ALL OK" if ok else "SOME FAILED", flush=True)

for L in [int(x) for x in sys.argv[1:]]:
    m = sd.XXZChain(L, nup=L // 2, ctx=ctx)
    for dtype in (np.float64,):
        x = m.vector(dtype).fill_seeded(1, 1e-4)
        y = m.vector(dtype)
        for _ in range(3):
            sd.apply_H_(y, x, m)
        ctx.sync(); dist.barrier()
        ctx.timer_start()
        for _ in range(10):
            sd.apply_H_(y, x, m)
        ms = allmax(ctx.timer_stop() / 10)
        if rank == 0:
            print(f"[world={world}] L={L} {np.dtype(dtype).name} path={m.info['kernel_path']}: {ms:.3f} ms/apply "
                  f"({16 * m.dim / ms / 1e6:.0f} GB/s algorithmic aggregate)", flush=True)
    del x, y, m
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
