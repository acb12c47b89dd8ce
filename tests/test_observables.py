"""Observables (Observables.jl:14-109, PublicAPI.jl:94-106) without a GPU: the oracle's restatement
against known answers, and the device kernel's per-state arithmetic (sd_obs.h: sum_i s_i s_{i+r} as
(L - 2 popc(state xor rot_r(state))) / 4), run on the CPU by tests/emul/emul_obs.cpp, against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle.oracle as orc
from conftest import ROOT

vp = ctypes.c_void_p


def emul():
    d = os.path.join(ROOT, "tests", "emul")
    subprocess.check_call(["make", "-C", d], stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(d, "libsd_emul_obs.so"))
    lib.emul_obs.argtypes = [ctypes.c_int, vp, ctypes.c_uint64, vp, ctypes.c_int, vp, vp]
    return lib


def test_oracle_observables_known_answers():
    # Neel product state: <S_i> = +-1/2 alternating, connected correlations vanish identically
    m = orc.XXZChain(6, nup=3)
    neel = orc.neel_state(m)
    assert np.allclose(orc.magnetization_per_site(neel, m), [0.5, -0.5, 0.5, -0.5, 0.5, -0.5], atol=1e-15)
    assert np.allclose(orc.connected_correlations(neel, m), 0.0, atol=1e-15)
    # Heisenberg ground state (a singlet): <S_i> = 0, C_0 = 1/4, S(q=0) = 0 (total Sz is sharp), S(q) real and >= 0
    m = orc.XXZChain(4, nup=2)
    _, gs = orc.groundstate(m, lanc_m=6, rng=np.random.default_rng(1))
    assert np.allclose(orc.magnetization_per_site(gs, m), 0.0, atol=1e-12)
    C = orc.connected_correlations(gs, m)
    assert abs(C[0] - 0.25) < 1e-12 and abs(C.sum()) < 1e-12
    Sq = orc.structure_factor_Sq(gs, m)
    assert len(Sq) == 4 and abs(Sq[0.0]) < 1e-12 and all(v > -1e-12 for v in Sq.values())
    assert Sq == orc.structure_factor(m, gs)                                   # test_PublicAPI.jl:135-150
    # nearest-neighbour correlation of the 4-site open chain from its energy: E0 = sum_bonds <S.S> = 3 <SzSz>_bonds
    sz = orc._sz_table(m)
    zz_nn = sum(float((gs ** 2) @ (sz[:, i] * sz[:, i + 1])) for i in range(3))
    assert abs(3 * zz_nn - (-1.616025403784439)) < 1e-10


@pytest.mark.parametrize("L,nup", [(6, 3), (9, 4), (12, 6), (7, None), (10, 0), (11, 11), (5, 1)])
@pytest.mark.parametrize("cplx", [False, True])
def test_kernel_arithmetic_matches_oracle(L, nup, cplx):
    lib = emul()
    rng = np.random.default_rng(L * 10 + (nup or 0))
    m = orc.XXZChain(L, nup=nup)
    N = len(m)
    psi = rng.standard_normal(N) + (1j * rng.standard_normal(N) if cplx else 0)
    if N > 7:
        psi[rng.integers(0, N, N // 7)] = 0.0                                 # zero weights are skipped
    psi = np.ascontiguousarray(psi / np.linalg.norm(psi))
    states = np.ascontiguousarray(m.states, dtype=np.uint64)
    mags, zz = np.zeros(L), np.zeros(L)
    flat = psi.view(np.float64) if cplx else psi
    assert lib.emul_obs(L, states.ctypes.data, N, flat.ctypes.data, 2 if cplx else 1, mags.ctypes.data, zz.ctypes.data) == 0
    assert np.allclose(mags, orc.magnetization_per_site(psi, m), atol=1e-13)
    C = np.array([(zz[r] - float(np.dot(mags, np.roll(mags, -r)))) / L for r in range(L)])
    assert np.allclose(C, orc.connected_correlations(psi, m), atol=1e-13)
