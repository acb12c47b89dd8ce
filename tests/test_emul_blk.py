"""CPU checks of the block-layout kernel (sd_blk.h, the headline H.psi kernel):
tests/emul/emul_blk.cpp runs the SAME __host__ __device__ functions the CUDA kernel
runs (tile header lanes, item body, fused epilogue), lane by lane, on block-layout
copies of the inputs and is compared with the oracle.  Covers the padded f64
pair-row / c128 row layout (bijection onto non-padding slots, padding stays zero),
tile bases and neighbour-tile pointers, prefix streams, tail / mid / crossing hops,
every epilogue mode with its reductions, and tile-aligned sharding over 1..8 ranks
(including ranks whose shard is empty).  The emulator is test infrastructure; the
product never loads it."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

import oracle.oracle as orc
from conftest import ROOT

vp = ctypes.c_void_p
ARGTYPES = ([ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, ctypes.c_uint64, vp, vp,
             ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
             vp, vp, vp, ctypes.c_double, ctypes.c_double, vp, vp, ctypes.c_uint64, vp, ctypes.c_int])


def load():
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.environ.get("SD_EMUL_BLK_LIB") or os.path.join(d, "libsd_emul_blk.so"))   # SD_EMUL_BLK_LIB: the ASan build (make asan)
    lib.emul_blk_apply.argtypes = ARGTYPES
    return lib


VARIANT = 0          # flags of emul_blk_apply: + 256 peers' shards visible only where the traffic plan says, + 512 remote-weighted shard bounds


@pytest.fixture(scope="module")
def emul():
    return load()


def P(a):
    return a.ctypes.data_as(vp) if a is not None else None


def model_lists(L, rng=None):
    if rng is None:
        return np.full(L - 1, 0.5), np.ones(L - 1), np.zeros(L)
    return rng.uniform(0.3, 1.5, L - 1), rng.uniform(-1, 1, L - 1), rng.uniform(-1, 1, L)


def oracle_model(L, k, Jhop, Jz, h):
    hop = [(i + 1, i + 2, Jhop[i]) for i in range(L - 1)]
    zz = [(i + 1, i + 2, Jz[i]) for i in range(L - 1)]
    return orc.build_model(L, nup=k, hopping=hop, onsite_field=h, zz=zz)


def oracle_apply(m, psi, NC):
    if NC == 1:
        ref = np.empty_like(psi)
        orc.apply_H_(ref, psi, m)
        return ref
    pc = psi.view(np.complex128).copy()
    rf = np.empty_like(pc)
    orc.apply_H_(rf, pc, m)
    return rf.view(np.float64).copy()


def run(lib, L, k, NC, world, states, psi, Jhop, Jz, h, mode=0, red=0, hscale=1.0, a=1.0, b=0.0,
        vprev=None, phi=None, acc=None, ck=0j, far_bytes=1 << 20, variant=None):
    variant = VARIANT if variant is None else variant
    N = len(states)
    out = np.full(N * NC, np.nan)
    redsum = np.zeros(4)
    bounds = np.zeros(world + 1, dtype=np.uint64)
    nstore = np.zeros(1, dtype=np.uint64)
    for r in range(world):
        redr = np.zeros(4)
        rc = lib.emul_blk_apply(L, k, P(Jhop), P(Jz), P(h), NC, P(states), N, P(psi), P(out), world, r,
                                mode, red, hscale, a, b, P(vprev), P(phi), P(acc), ck.real, ck.imag,
                                P(redr), P(bounds), far_bytes, P(nstore), variant)
        assert rc == 0, rc
        redsum += redr
    return out, redsum, bounds, int(nstore[0])


CASES = [(16, 8), (16, 2), (16, 14), (16, 0), (16, 16), (16, 1), (16, 15), (17, 8), (17, 3), (18, 9), (18, 12),
         (19, 9), (20, 10), (20, 4)]


@pytest.mark.parametrize("L,k", CASES)
@pytest.mark.parametrize("NC", [1, 2])
def test_block_body_matches_oracle(emul, L, k, NC):
    rng = np.random.default_rng(L * 1000 + k * 10 + NC)
    Jhop, Jz, h = model_lists(L, rng)
    m = oracle_model(L, k, Jhop, Jz, h)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    N = len(states)
    psi = rng.standard_normal(N * NC)
    ref = oracle_apply(m, psi, NC)
    worlds = (1, 2, 3, 8) if L <= 18 else (1, 4)
    for world in worlds:
        out, _, bounds, nstore = run(emul, L, k, NC, world, states, psi, Jhop, Jz, h)
        assert np.linalg.norm(out - ref) <= 1e-14 * max(1.0, np.linalg.norm(ref)), world
        assert bounds[0] == 0 and bounds[-1] == N and np.all(np.diff(bounds.astype(np.int64)) >= 0)
        assert N <= nstore <= 1.25 * N + 64 * (1 << (L - 15))          # padding overhead stays small


@pytest.mark.parametrize("L,k", [(22, 11), (24, 12), (24, 9)])
def test_block_body_longer_prefix(emul, L, k):
    """L = 22, 24 (7 / 9 prefix sites: up to 8 prefix-bond partner tiles + the crossing partner per tile), f64, 1 and 5 ranks."""
    rng = np.random.default_rng(L)
    Jhop, Jz, h = model_lists(L, rng)
    m = oracle_model(L, k, Jhop, Jz, h)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    psi = rng.standard_normal(len(states))
    ref = oracle_apply(m, psi, 1)
    for world in (1, 5):
        out, _, _, _ = run(emul, L, k, 1, world, states, psi, Jhop, Jz, h)
        assert np.linalg.norm(out - ref) <= 1e-14 * np.linalg.norm(ref), world


def test_zero_couplings_on_some_bonds(emul):
    """J = 0 on a prefix bond, a mid bond, the crossing bonds and a tail bond: inactive bonds are skipped
    in the header / item tables, not multiplied by zero."""
    L, k = 18, 9
    rng = np.random.default_rng(5)
    Jhop, Jz, h = model_lists(L, rng)
    for p in (1, 2, 6, 12, 14, 16):
        Jhop[p] = 0.0
    m = oracle_model(L, k, Jhop, Jz, h)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    for NC in (1, 2) if VARIANT < 2 else (1,):
        psi = rng.standard_normal(len(states) * NC)
        ref = oracle_apply(m, psi, NC)
        out, _, _, _ = run(emul, L, k, NC, 2, states, psi, Jhop, Jz, h)
        assert np.linalg.norm(out - ref) <= 1e-14 * np.linalg.norm(ref)


@pytest.mark.parametrize("far_bytes", [0, 1 << 12, 1 << 40])
def test_far_near_sorting_of_neighbour_tiles_is_only_an_order(emul, far_bytes):
    L, k = 19, 10
    rng = np.random.default_rng(2)
    Jhop, Jz, h = model_lists(L, rng)
    m = oracle_model(L, k, Jhop, Jz, h)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    psi = rng.standard_normal(len(states))
    ref = oracle_apply(m, psi, 1)
    out, _, _, _ = run(emul, L, k, 1, 1, states, psi, Jhop, Jz, h, far_bytes=far_bytes)
    assert np.linalg.norm(out - ref) <= 1e-14 * np.linalg.norm(ref)


@pytest.mark.parametrize("NC", [1, 2])
@pytest.mark.parametrize("L,k", [(16, 8), (17, 9), (18, 7)])
def test_fused_epilogues(emul, L, k, NC):
    """Every epilogue the recurrences use (Hamiltonian.jl:286-301, KPM_Sqw.jl:111-117,
    Chebyshev.jl:112-116, Lanczos.jl:50), sharded over 2 ranks."""
    rng = np.random.default_rng(4 + L)
    Jhop, Jz, h = model_lists(L, rng)
    m = oracle_model(L, k, Jhop, Jz, h)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    N = len(states)
    v, vprev, phi, acc0 = (rng.standard_normal(N * NC) for _ in range(4))
    a, b, hs = 2.5, 0.3, -1.0
    ck = (0.4 - 0.7j) if NC == 2 else (0.4 + 0j)
    Hv = oracle_apply(m, v, NC)
    cplx = (lambda x: x.view(np.complex128)) if NC == 2 else (lambda x: x)
    # apply + <psi, H psi> (Lanczos alpha)
    out, red, _, _ = run(emul, L, k, NC, 2, states, v, Jhop, Jz, h, red=1)
    assert np.linalg.norm(out - Hv) <= 1e-14 * np.linalg.norm(Hv)
    d = np.vdot(cplx(v), cplx(Hv))
    assert abs(complex(red[0], red[1]) - d) < 1e-10 * max(1.0, abs(d))
    # rescaled
    resc = (Hv - b * v) / a
    out, _, _, _ = run(emul, L, k, NC, 2, states, v, Jhop, Jz, h, mode=1, a=a, b=b)
    assert np.linalg.norm(out - resc) <= 1e-14 * np.linalg.norm(resc)
    # Chebyshev step of -H with accumulation and all reductions
    nxt = 2.0 * ((hs * Hv - b * v) / a) - vprev
    acc = acc0.copy()
    out, red, _, _ = run(emul, L, k, NC, 2, states, v, Jhop, Jz, h, mode=2, red=7, hscale=hs, a=a, b=b,
                         vprev=vprev, phi=phi, acc=acc, ck=ck)
    assert np.linalg.norm(out - nxt) <= 1e-14 * np.linalg.norm(nxt)
    assert abs(complex(red[0], red[1]) - np.vdot(cplx(v), cplx(nxt))) < 1e-10
    assert abs(red[2] - np.vdot(cplx(phi), cplx(nxt)).real) < 1e-10
    assert abs(red[3] - nxt @ nxt) < 1e-9
    want = cplx(acc0) + (ck if NC == 2 else ck.real) * cplx(nxt)
    assert np.linalg.norm(cplx(acc) - want) < 1e-12


def _gloo_worker(rank, world, port, q):
    """world_size-2 gloo run of the sharded block path's host logic: every rank runs the emulated
    kernel on its own tile-aligned shard (peer shards are separate buffers selected by the owner
    lookup in the tile header), the shards are gathered and the fused dot is all-reduced."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = load()
        L, k = 18, 9
        Jhop, Jz, h = model_lists(L)
        m = oracle_model(L, k, Jhop, Jz, h)
        states = np.ascontiguousarray(m.states, dtype=np.uint64)
        N = len(states)
        psi = orc.fill_seeded(N, 99)
        out = np.zeros(N)
        red = np.zeros(4)
        bounds = np.zeros(world + 1, dtype=np.uint64)
        rc = lib.emul_blk_apply(L, k, P(Jhop), P(Jz), P(h), 1, P(states), N, P(psi), P(out), world, rank,
                                0, 1, 1.0, 1.0, 0.0, None, None, None, 0.0, 0.0, P(red), P(bounds), 1 << 20, None, 0)
        assert rc == 0
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        sizes = [int(bounds[g + 1] - bounds[g]) for g in range(world)]
        parts = [torch.zeros(s, dtype=torch.float64) for s in sizes]
        for g in range(world):                              # ragged shards: gather by broadcast
            if g == rank:
                parts[g] = torch.from_numpy(out[lo:hi].copy())
            dist.broadcast(parts[g], src=g)
        full = torch.cat(parts).numpy()
        dot = torch.tensor([red[0]], dtype=torch.float64)
        dist.all_reduce(dot)                                # the scalar all-reduce of the Lanczos alpha
        if rank == 0:
            ref = oracle_apply(m, psi, 1)
            q.put((float(np.linalg.norm(full - ref) / np.linalg.norm(ref)), float(abs(dot.item() - psi @ ref)), sizes))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_block_apply(emul):
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
    err, derr, sizes = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-14 and derr < 1e-9 and sum(sizes) == orc.lib().orc_sector_dim(18, 9) and min(sizes) > 0


def test_block_body_matches_golden_fixture(emul):
    """tests/golden/apply_golden.npz holds oracle outputs for the seeded psi at the block kernel's
    smallest sizes (L = 16, 17); the GPU golden test reads the same file."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "apply_golden.npz"))
    seen = 0
    for key in [k for k in g.files if k.startswith("out_")]:
        _, L, nup, kind = key.split("_")
        L, nup = int(L), int(nup)
        if L < 16:
            continue
        seen += 1
        NC = 2 if kind == "c128" else 1
        if VARIANT >= 2 and NC == 2:
            continue
        Jhop, Jz, h = np.full(L - 1, float(g["Jxy"]) / 2), np.full(L - 1, float(g["Jz"])), np.full(L, float(g["hz"]))
        om = orc.XXZChain(L, nup=nup)
        states = np.array(om.states, dtype=np.uint64)
        psi = orc.fill_seeded(len(states), int(g["seed"]), cplx=NC == 2).view(np.float64).copy()
        out, _, _, _ = run(emul, L, nup, NC, 2, states, psi, Jhop, Jz, h)
        ref = np.ascontiguousarray(g[key]).view(np.float64)
        assert np.linalg.norm(out - ref) <= 1e-14 * np.linalg.norm(ref), key
    assert seen == 4


@pytest.mark.parametrize("L,k,N", [(32, 16, 601080390), (34, 17, 2333606220), (36, 18, 9075135300), (32, 10, 64512240)])
def test_shard_plan_at_full_sizes(emul, L, k, N):
    """Host-side layout/shard plan at BASELINE.json's sizes (L=36: ranks and stored offsets beyond 2^32):
    bounds are tile aligned and monotone, shards are balanced, padding overhead is small, and the
    per-rank stored ranges tile the stored vector exactly."""
    emul.emul_blk_plan.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp]
    for world in (1, 2, 4, 8):
        bounds, pstart, keys = (np.zeros(world + 1, dtype=np.uint64) for _ in range(3))
        nstore, ntiles, cap = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint32)
        assert emul.emul_blk_plan(L, k, world, P(bounds), P(pstart), P(keys), P(nstore), P(cap), P(ntiles)) == 0
        b, ps, ks = bounds.astype(object), pstart.astype(object), keys.astype(object)
        assert b[0] == 0 and b[-1] == N and ps[0] == 0 and ps[-1] == int(nstore[0])
        assert 0 <= ks[0] and ks[-1] <= 1 << (L - 15)                  # the first / last keys may be impossible prefixes
        assert all(b[g] < b[g + 1] and ps[g] < ps[g + 1] and ks[g] < ks[g + 1] for g in range(world))
        sizes = [b[g + 1] - b[g] for g in range(world)]
        assert max(sizes) - min(sizes) <= 2 * 6435                   # a rank boundary moves by at most one tile
        slack = 1.03 if 2 * k == L else 1.15                          # zero padding of the block layout (1.1-1.2 % at Sz = 0)
        assert N <= int(nstore[0]) <= (slack if 2 * k == L else 1.10) * N
        assert all(p % 16 == 0 for p in ps)                           # 128-byte aligned shard bases (TMA needs 16)
        assert int(cap[0]) == 6480 and int(cap[0]) * 8 * 3 < 227 * 1024
        stored = [ps[g + 1] - ps[g] for g in range(world)]
        assert max(stored) <= slack * max(sizes) + 6480


@pytest.mark.parametrize("L,k,e,world", [(24, 12, 5, 1), (26, 13, 12, 1), (28, 14, 12, 4), (20, 10, 3, 2), (26, 13, 10, 1)])
def test_tile_order_is_a_permutation_of_the_shards_valid_tiles(emul, L, k, e, world):
    """sd_blk_tile_order (breadth-first order of the popcount groups, the default for vectors beyond the L2): every valid
    tile key of the shard exactly once; equal to the python model of the same order that scripts/l2_sim.py evaluates."""
    emul.emul_blk_order.argtypes = [ctypes.c_int] * 5 + [vp, ctypes.c_long, vp, vp]
    emul.emul_blk_order.restype = ctypes.c_long
    A = L - 15
    allkeys = []
    for rank in range(world):
        out = np.zeros(1 << A, dtype=np.uint32)
        lo, hi = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64)
        n = emul.emul_blk_order(L, k, world, rank, e, P(out), len(out), P(lo), P(hi))
        assert n >= 0
        keys = out[:n].astype(np.int64)
        valid = [key for key in range(int(lo[0]), int(hi[0]))
                 if 0 <= k - (A - bin(key).count("1")) <= 15]          # key bit = NOT prefix bit
        assert sorted(keys.tolist()) == valid
        allkeys.append(keys)
    if world == 1:
        import importlib.util
        import sys
        argv = sys.argv
        sys.argv = ["l2_sim", str(L)]
        try:
            spec = importlib.util.spec_from_file_location("l2_sim", os.path.join(ROOT, "scripts", "l2_sim.py"))
            sim = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(sim)
        finally:
            sys.argv = argv
        ref = sim.bfs_order(e)                                          # prefix bit patterns
        ref_keys = [sum((0 if (Pb >> q) & 1 else 1) << (A - 1 - q) for q in range(A)) for Pb in ref]
        assert ref_keys == allkeys[0].tolist()


@pytest.mark.parametrize("L,k,world,chunks", [(20, 10, 2, 4), (22, 11, 3, 5), (24, 12, 4, 8), (26, 9, 8, 3), (28, 14, 8, 8), (32, 16, 8, 8),
                                              (32, 16, 2, 8), (36, 18, 8, 16)])
def test_remote_volume_plan(L, k, world, chunks):
    """sd_shard_host.h: for every rank, the peer ranges its tile headers point at (same header code as the kernel), listed
    chunk by chunk: they cover every remote partner tile of chunks 0..j, segments are disjoint, tile aligned and inside
    their peer's shard; a handful of large segments per rank.  This is the traffic model behind the weighted shards.  At half filling the
    busiest rank pulls 0.5 shards at 2 ranks, 1.5 at 4 and 2.5 at 8 (the ranks whose top prefix bits are 101 / 010 have
    both top bonds active plus half of the third), the average over ranks is about 0.5 / 1.0 / 1.5."""
    lib = load()
    lib.emul_halo_plan.argtypes = [ctypes.c_int] * 5 + [vp]
    ratios = []
    for rank in (range(world) if L <= 28 else sorted({0, world // 2, world - 1})):
        st = np.zeros(8, dtype=np.uint64)
        assert lib.emul_halo_plan(L, k, world, rank, chunks, P(st)) == 0, rank
        nseg, remote, local, peers, maxseg = (int(x) for x in st[:5])
        assert local > 0 and 1 <= peers <= world - 1 and nseg <= 8 * chunks and maxseg <= 12
        ratios.append(remote / local)
    if 2 * k == L and world in (2, 4, 8) and L <= 28:
        want_max, want_avg = {2: (0.5, 0.5), 4: (1.5, 1.0), 8: (2.5, 1.5)}[world]
        assert abs(max(ratios) - want_max) < 0.1 and abs(np.mean(ratios) - want_avg) < 0.1


@pytest.mark.parametrize("L,k,world", [(24, 12, 8), (28, 14, 8), (28, 14, 4), (26, 10, 8)])
def test_remote_weighted_shard_bounds(L, k, world):
    """sd_shard_balance (the default for more than two ranks): cut positions weighted by remote volume.  Bounds stay tile aligned and
    monotone, no rank gets slower than the slowest rank of the equal split, and at 8 ranks of a half-filled chain the
    largest per-rank time max(local, 0.7 * remote) drops by about a third (1.77 -> 1.19 shares: the 101 / 010 ranks
    get half-size shards)."""
    lib = load()
    lib.emul_halo_balance.argtypes = [ctypes.c_int] * 3 + [ctypes.c_double, ctypes.c_int, vp, vp, vp, vp]
    b, kk = np.zeros(world + 1, dtype=np.uint64), np.zeros(world + 1, dtype=np.uint64)
    cost, per = np.zeros(2), np.zeros(world)
    assert lib.emul_halo_balance(L, k, world, 0.7, 8, P(b), P(kk), P(cost), P(per)) == 0
    N = math.comb(L, k)
    bb, kb = b.astype(object), kk.astype(object)
    assert bb[0] == 0 and bb[-1] == N and all(bb[g] < bb[g + 1] and kb[g] < kb[g + 1] for g in range(world))
    assert cost[1] <= cost[0] * (1 + 1e-12) and abs(per.max() - cost[1]) <= 1e-9 * cost[1]
    share = N / world
    if 2 * k == L and world == 8:
        assert 1.7 < cost[0] / share < 1.85 and cost[1] / share < 1.3
        sizes = np.diff(b.astype(np.float64)) / share
        assert sizes[2] < 0.7 and sizes[5] < 0.7 and sizes[0] > 1.05


@pytest.mark.parametrize("variant", [256, 256 + 512, 512], ids=["plan", "plan_bal", "bal"])
@pytest.mark.parametrize("L,k,world", [(18, 9, 2), (20, 10, 4), (20, 10, 8), (22, 11, 8), (20, 6, 5)])
def test_sharded_apply_with_planned_peer_ranges_and_weighted_shards(variant, L, k, world):
    """The traffic plan and the weighted shards end to end on the CPU: every rank runs the emulated kernel on NaN-filled copies of
    its peers' shards that hold only what the plan lists, chunk by chunk, before that chunk's tiles run -- a partner
    tile the plan missed would put NaN into the result -- with equal and with remote-weighted shard bounds, including the
    fused epilogue with all reductions."""
    lib = load()
    rng = np.random.default_rng(L + world)
    Jhop, Jz, h = model_lists(L, rng)
    m = oracle_model(L, k, Jhop, Jz, h)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    N = len(states)
    for NC in (1, 2):
        psi, vprev, phi = (rng.standard_normal(N * NC) for _ in range(3))
        ref = oracle_apply(m, psi, NC)
        out, _, bounds, _ = run(lib, L, k, NC, world, states, psi, Jhop, Jz, h, variant=variant)
        assert np.linalg.norm(out - ref) <= 1e-14 * np.linalg.norm(ref)
        assert bounds[0] == 0 and bounds[-1] == N and np.all(np.diff(bounds.astype(np.int64)) >= 0)
        a, b = 2.5, 0.3
        nxt = 2.0 * ((ref - b * psi) / a) - vprev
        out, red, _, _ = run(lib, L, k, NC, world, states, psi, Jhop, Jz, h, mode=2, red=7, a=a, b=b, vprev=vprev, phi=phi, variant=variant)
        cplx = (lambda x: x.view(np.complex128)) if NC == 2 else (lambda x: x)
        assert np.linalg.norm(out - nxt) <= 1e-14 * np.linalg.norm(nxt)
        assert abs(complex(red[0], red[1]) - np.vdot(cplx(psi), cplx(nxt))) < 1e-9
        assert abs(red[2] - np.vdot(cplx(phi), cplx(nxt)).real) < 1e-9 and abs(red[3] - nxt @ nxt) < 1e-8



@pytest.mark.parametrize("NC", [1, 2])
@pytest.mark.parametrize("L,k", [(16, 8), (16, 3), (17, 9), (18, 9), (20, 10), (16, 16), (16, 1)])
def test_periodic_wrap_bond_variant(emul, L, k, NC):
    """XXZChain(boundary=:periodic) (SpinModel.jl:71-78) on the block path: the WRAP variant of the item body (the tile
    header carries the wrap partner tile; the bond between tail site T-1 and prefix site 0 is one more element read per
    tail configuration), every epilogue, 1 .. 3 ranks -- against the oracle's apply_H! on the model with the (L, 1) bonds."""
    rng = np.random.default_rng(500 + L * 10 + k + NC)
    Jhop, Jz, h = model_lists(L, rng)
    Jw, Jzw = float(rng.uniform(0.3, 1.5)), float(rng.uniform(-1, 1))
    hop = [(i + 1, i + 2, Jhop[i]) for i in range(L - 1)] + [(L, 1, Jw)]
    zz = [(i + 1, i + 2, Jz[i]) for i in range(L - 1)] + [(L, 1, Jzw)]
    m = orc.build_model(L, nup=k, hopping=hop, onsite_field=h, zz=zz)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    N = len(states)
    psi = rng.standard_normal(N * NC)
    ref = oracle_apply(m, psi, NC)
    emul.emul_blk_set_wrap.argtypes = [ctypes.c_double, ctypes.c_double]
    emul.emul_blk_set_wrap.restype = None
    emul.emul_blk_set_wrap(Jw, Jzw)
    try:
        for world in (1, 2, 3):
            out, _, _, _ = run(emul, L, k, NC, world, states, psi, Jhop, Jz, h)
            assert np.linalg.norm(out - ref) <= 1e-14 * max(1.0, np.linalg.norm(ref)), world
        # fused epilogue on top of the add-in: (H psi - b psi) / a with the dot <psi, out>
        a, b = 3.0, -0.4
        out, red, _, _ = run(emul, L, k, NC, 1, states, psi, Jhop, Jz, h, mode=1, red=1, a=a, b=b)
        want = (ref - b * psi) / a
        assert np.linalg.norm(out - want) <= 1e-14 * max(1.0, np.linalg.norm(want))
        if NC == 1:
            assert abs(red[0] - float(psi @ want)) <= 1e-12 * max(1.0, abs(float(psi @ want)))
    finally:
        emul.emul_blk_set_wrap(0.0, 0.0)
