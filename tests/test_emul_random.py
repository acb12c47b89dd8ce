"""Randomised sweep of the CPU emulation of the block kernel's item body (sd_blkl.h):
random chain length and filling, random per-bond couplings with some set to zero, random number of ranks, random fused
epilogue and reductions, f64 and c128 -- each case against the oracle."""
import numpy as np
import pytest

import test_emul_blk as T


@pytest.mark.parametrize("seed", range(10))
def test_random_models_shards_and_epilogues(seed):
    lib = T.load()
    rng = np.random.default_rng(1000 + seed)
    L = int(rng.integers(16, 20))
    k = int(rng.integers(0, L + 1))
    Jhop, Jz, h = T.model_lists(L, rng)
    for p in range(L - 1):
        if rng.random() < 0.15:
            Jhop[p] = 0.0
        if rng.random() < 0.1:
            Jz[p] = 0.0
    m = T.oracle_model(L, k, Jhop, Jz, h)
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    N = len(states)
    world = int(rng.integers(1, 9))
    for NC in (1, 2):
        psi = rng.standard_normal(N * NC)
        ref = T.oracle_apply(m, psi, NC)
        cpl = (lambda x: x.view(np.complex128)) if NC == 2 else (lambda x: x)
        for variant in (0,):
            mode = int(rng.integers(0, 3))
            red = int(rng.integers(0, 8)) if mode == 2 else int(rng.integers(0, 2))
            a, b, hs = 2.5, 0.3, (-1.0 if rng.random() < 0.3 else 1.0)
            vprev = rng.standard_normal(N * NC) if mode == 2 else None
            phi = rng.standard_normal(N * NC) if (red & 2) else None
            out, rs, _, _ = T.run(lib, L, k, NC, world, states, psi, Jhop, Jz, h, mode=mode, red=red, hscale=hs, a=a, b=b,
                                  vprev=vprev, phi=phi, variant=variant, far_bytes=int(rng.choice([0, 1 << 12, 1 << 20, 1 << 40])))
            want = hs * ref
            if mode >= 1:
                want = (want - b * psi) / a
            if mode == 2:
                want = 2 * want - vprev
            tag = (L, k, NC, world, variant, mode, red)
            assert np.linalg.norm(out - want) <= 1e-13 * max(1.0, np.linalg.norm(want)), tag
            if red & 1:
                d = np.vdot(cpl(psi), cpl(want))
                assert abs(complex(rs[0], rs[1]) - d) < 1e-9 * max(1, abs(d)), tag
            if red & 2:
                assert abs(rs[2] - np.vdot(cpl(phi), cpl(want)).real) < 1e-9 * max(1, np.linalg.norm(want) * np.linalg.norm(phi)), tag
            if red & 4:
                assert abs(rs[3] - want @ want) < 1e-9 * max(1, want @ want), tag
