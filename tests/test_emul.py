"""CPU checks of the tiled kernel's index algebra: tests/emul/emul.cpp runs the
SAME __host__ __device__ phase functions the CUDA kernel runs (sd_tile.h),
sequentially, and is compared with the oracle.  This validates tile bases,
neighbour-tile shifts, the class-major shared-memory permutation, tail / mid /
crossing hops, the fused epilogue and the shard-ownership logic without a GPU.
The emulator is test infrastructure; the product never loads it."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle.oracle as orc
from conftest import ROOT

vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def emul():
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(d, "libsd_emul.so"))
    lib.emul_tile_apply.argtypes = ([ctypes.c_int] * 4 + [vp, vp, vp, ctypes.c_int, vp, vp, ctypes.c_uint,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_double, ctypes.c_double, ctypes.c_double, vp, vp, vp,
                                    ctypes.c_double, ctypes.c_double, vp, vp, ctypes.c_uint64])
    return lib


def P(a):
    return a.ctypes.data_as(vp) if a is not None else None


def model_lists(L, rng=None):
    if rng is None:
        return np.full(L - 1, 0.5), np.ones(L - 1), np.zeros(L)
    return rng.uniform(0.3, 1.5, L - 1), rng.uniform(-1, 1, L - 1), rng.uniform(-1, 1, L)


def oracle_apply(L, k, Jhop, Jz, h, psi, NC):
    hop = [(i + 1, i + 2, Jhop[i]) for i in range(L - 1)]
    zz = [(i + 1, i + 2, Jz[i]) for i in range(L - 1)]
    m = orc.build_model(L, nup=k, hopping=hop, onsite_field=h, zz=zz)
    if NC == 1:
        ref = np.empty_like(psi)
        orc.apply_H_(ref, psi, m)
        return ref
    pc = psi.view(np.complex128).copy()
    rf = np.empty_like(pc)
    orc.apply_H_(rf, pc, m)
    return rf.view(np.float64).copy()


def run(lib, L, k, B, T, NC, world, psi, Jhop, Jz, h, nthreads=64, mode=0, red=0, hscale=1.0, a=1.0, b=0.0,
        vprev=None, phi=None, acc=None, ck=0j, far=50):
    N = len(psi) // NC
    out = np.full(N * NC, np.nan)
    redsum = np.zeros(4)
    bounds = np.zeros(world + 1, dtype=np.uint64)
    for r in range(world):
        redr = np.zeros(4)
        rc = lib.emul_tile_apply(L, k, B, T, P(Jhop), P(Jz), P(h), NC, P(psi), P(out), nthreads, world, r,
                                 mode, red, hscale, a, b, P(vprev), P(phi), P(acc), ck.real, ck.imag, P(redr), P(bounds), far)
        assert rc == 0
        redsum += redr
    return out, redsum, bounds


CASES = [(10, 5, 8, 5), (12, 6, 9, 5), (12, 4, 10, 5), (14, 7, 10, 5), (14, 7, 12, 4), (16, 8, 12, 5), (13, 6, 13, 5),
         (12, 0, 9, 5), (12, 12, 9, 5), (12, 1, 9, 5), (12, 11, 9, 5), (16, 8, 14, 6), (12, 6, 8, 3), (16, 8, 15, 5),
         (15, 7, 13, 4)]


@pytest.mark.parametrize("L,k,B,T", CASES)
@pytest.mark.parametrize("NC", [1, 2])
def test_tile_body_matches_oracle(emul, L, k, B, T, NC):
    if NC == 2 and T in (3, 6):
        pytest.skip("combination not instantiated in the emulator")
    rng = np.random.default_rng(L * 1000 + k * 10 + B)
    Jhop, Jz, h = model_lists(L, rng)
    N = orc.lib().orc_sector_dim(L, k)
    psi = rng.standard_normal(N * NC)
    ref = oracle_apply(L, k, Jhop, Jz, h, psi, NC)
    for world in (1, 2, 3):
        out, _, bounds = run(emul, L, k, B, T, NC, world, psi, Jhop, Jz, h)
        assert np.linalg.norm(out - ref) <= 1e-14 * max(1.0, np.linalg.norm(ref)), world
        assert bounds[0] == 0 and bounds[-1] == N and np.all(np.diff(bounds.astype(np.int64)) >= 0)


@pytest.mark.parametrize("nthreads", [32, 96, 512])
def test_tile_body_independent_of_cta_size(emul, nthreads):
    L, k, B, T = 14, 7, 11, 5
    rng = np.random.default_rng(0)
    Jhop, Jz, h = model_lists(L)
    psi = rng.standard_normal(orc.lib().orc_sector_dim(L, k))
    ref = oracle_apply(L, k, Jhop, Jz, h, psi, 1)
    out, _, _ = run(emul, L, k, B, T, 1, 1, psi, Jhop, Jz, h, nthreads=nthreads)
    assert np.linalg.norm(out - ref) <= 1e-14 * np.linalg.norm(ref)


@pytest.mark.parametrize("NC", [1, 2])
def test_fused_epilogue(emul, NC):
    """mode 2 (Chebyshev step) + acc + all reductions, sharded over 2 ranks."""
    L, k, B, T = 14, 6, 10, 5
    rng = np.random.default_rng(4)
    Jhop, Jz, h = model_lists(L, rng)
    N = orc.lib().orc_sector_dim(L, k)
    v, vprev, phi, acc0 = (rng.standard_normal(N * NC) for _ in range(4))
    a, b, hs = 2.5, 0.3, -1.0
    ck = (0.4 - 0.7j) if NC == 2 else (0.4 + 0j)
    Hv = oracle_apply(L, k, Jhop, Jz, h, v, NC)
    nxt = 2.0 * ((hs * Hv - b * v) / a) - vprev
    acc = acc0.copy()
    out, red, _ = run(emul, L, k, B, T, NC, 2, v, Jhop, Jz, h, mode=2, red=7, hscale=hs, a=a, b=b, vprev=vprev,
                      phi=phi, acc=acc, ck=ck)
    assert np.linalg.norm(out - nxt) <= 1e-14 * np.linalg.norm(nxt)
    if NC == 2:
        vc, nc_, pc = v.view(np.complex128), nxt.view(np.complex128), phi.view(np.complex128)
        assert abs(complex(red[0], red[1]) - np.vdot(vc, nc_)) < 1e-10
        assert abs(red[2] - np.vdot(pc, nc_).real) < 1e-10
        assert np.linalg.norm(acc.view(np.complex128) - (acc0.view(np.complex128) + ck * nc_)) < 1e-12
    else:
        assert abs(red[0] - v @ nxt) < 1e-10 and abs(red[2] - phi @ nxt) < 1e-10
        assert np.linalg.norm(acc - (acc0 + ck.real * nxt)) < 1e-12
    assert abs(red[3] - nxt @ nxt) < 1e-9


def _gloo_worker(rank, world, port, q):
    """world_size-2 gloo run of the N>1 path's host logic: every rank computes
    its own shard with the emulated kernel (peer shards are separate buffers
    selected by the owner lookup) and the shards are all-gathered."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = os.path.join(ROOT, "tests", "emul")
        lib = ctypes.CDLL(os.path.join(d, "libsd_emul.so"))
        lib.emul_tile_apply.argtypes = ([ctypes.c_int] * 4 + [vp, vp, vp, ctypes.c_int, vp, vp, ctypes.c_uint,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_double, ctypes.c_double, ctypes.c_double, vp, vp, vp,
                                        ctypes.c_double, ctypes.c_double, vp, vp, ctypes.c_uint64])
        L, k, B, T = 16, 8, 11, 5
        Jhop, Jz, h = model_lists(L)
        N = orc.lib().orc_sector_dim(L, k)
        psi = orc.fill_seeded(N, 99)
        out = np.zeros(N)
        red = np.zeros(4)
        bounds = np.zeros(world + 1, dtype=np.uint64)
        rc = lib.emul_tile_apply(L, k, B, T, P(Jhop), P(Jz), P(h), 1, P(psi), P(out), 128, world, rank,
                                 0, 1, 1.0, 1.0, 0.0, None, None, None, 0.0, 0.0, P(red), P(bounds), 100)
        assert rc == 0
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        mine = torch.from_numpy(out[lo:hi].copy())
        sizes = [int(bounds[g + 1] - bounds[g]) for g in range(world)]
        parts = [torch.zeros(s, dtype=torch.float64) for s in sizes]
        dist.all_gather(parts, mine) if len(set(sizes)) == 1 else None
        if len(set(sizes)) != 1:                        # ragged shards: gather by broadcast
            for g in range(world):
                parts[g] = mine.clone() if g == rank else parts[g]
                dist.broadcast(parts[g], src=g)
        full = torch.cat(parts).numpy()
        dot = torch.tensor([red[0]], dtype=torch.float64)
        dist.all_reduce(dot)                            # the scalar all-reduce of the Lanczos alpha
        if rank == 0:
            ref = oracle_apply(L, k, Jhop, Jz, h, psi, 1)
            q.put((float(np.linalg.norm(full - ref) / np.linalg.norm(ref)), float(abs(dot.item() - psi @ ref)),
                   sizes))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_apply(emul):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, derr, sizes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-14 and derr < 1e-9 and sum(sizes) == orc.lib().orc_sector_dim(16, 8)
