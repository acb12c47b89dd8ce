"""GPU checks of the memory-lean ground state (sd_lanczos_lean; SURVEY.md 8f-3, an extension): E0 within 1e-10 of
the reference-faithful device path and of the ARPACK/oracle golden energies, Ritz vector normalised with a small
residual, breakdown handling, bit-identical repeat."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.oracle as orc  # noqa: E402
from conftest import sd  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "energy_golden.json")))


@pytest.mark.parametrize("L,nup,lanc_m", [(10, 5, 60), (12, 4, 200), (16, 8, 150), (18, 9, 150)])
def test_lean_equals_faithful(L, nup, lanc_m):
    m = sd.XXZChain(L, nup=nup)
    v0 = np.random.default_rng(L).standard_normal(m.dim)
    E, psi = sd.lanczos_groundstate_lean(sd.apply_H_, m, lanc_m=lanc_m, v0=v0)
    Eref, pref = sd.lanczos_groundstate(sd.apply_H_, m, lanc_m=lanc_m, v0=v0)
    assert abs(E - Eref) < 1e-10
    assert abs(np.linalg.norm(psi) - 1) < 1e-12
    assert min(np.linalg.norm(psi - pref), np.linalg.norm(psi + pref)) < 1e-6
    E2, psi2 = sd.groundstate(m, method="lanczos_lean", lanc_m=lanc_m, v0=v0)
    assert E2 == E and np.array_equal(psi, psi2)                            # deterministic kernels


@pytest.mark.parametrize("L", [20, 24])
def test_lean_against_golden_energy_device_resident(L):
    m = sd.XXZChain(L, nup=L // 2)
    v0 = m.vector(np.float64).fill_seeded(3)
    E0, psi = sd.lanczos_groundstate_lean(sd.apply_H_, m, lanc_m=160, v0=v0, device=True)
    assert abs(E0 - GOLD["E0"][str(L)]) < 1e-10
    h = m.vector(np.float64)
    sd.apply_H_(h, psi, m)
    assert abs(psi.dot(h).real - E0) < 1e-10
    h.axpy(-E0, psi)
    assert h.norm() < 1e-6


def test_lean_breakdown():
    m, om = sd.XXZChain(8, nup=4), orc.XXZChain(8, nup=4)
    _, gs = orc.lanczos_groundstate(orc.apply_H_, om, lanc_m=70, v0=np.random.default_rng(0).standard_normal(len(om)))
    E, psi, a, b = sd.lanczos_groundstate_lean(sd.apply_H_, m, lanc_m=40, tol=1e-8, v0=gs, return_tridiag=True)
    assert len(a) < 40 and abs(E + 3.374932598687896) < 1e-10 and abs(abs(psi @ gs) - 1) < 1e-10
    with pytest.raises(sd.ZeroNormError):
        sd.lanczos_groundstate_lean(sd.apply_H_, m, v0=np.zeros(m.dim))


@pytest.mark.parametrize("L,nup,lanc_m", [(10, 5, 80), (16, 8, 100), (12, 6, 150)])
def test_batched_check_pass_equals_one_at_a_time(monkeypatch, L, nup, lanc_m):
    """SD_BATCH_CHECK=1 (sd_bdot.cuh): the check pass of the full reorthogonalisation (Lanczos.jl:142-153) with the
    overlaps taken eight at a time and fetched once -- same E0, same Ritz vector, same tridiagonal matrix as the
    one-at-a-time loop (the two differ only in the summation order of the final reduction of each overlap)."""
    m = sd.XXZChain(L, nup=nup)
    v0 = np.random.default_rng(L).standard_normal(m.dim)
    monkeypatch.delenv("SD_BATCH_CHECK", raising=False)
    E_a, psi_a, al_a, be_a = sd.lanczos_groundstate(sd.apply_H_, m, lanc_m=lanc_m, v0=v0, return_tridiag=True)
    monkeypatch.setenv("SD_BATCH_CHECK", "1")
    E_b, psi_b, al_b, be_b = sd.lanczos_groundstate(sd.apply_H_, m, lanc_m=lanc_m, v0=v0, return_tridiag=True)
    assert abs(E_a - E_b) < 1e-12 and len(al_a) == len(al_b)
    assert np.allclose(al_a, al_b, atol=1e-9) and np.allclose(be_a, be_b, atol=1e-9)
    assert min(np.linalg.norm(psi_a - psi_b), np.linalg.norm(psi_a + psi_b)) < 1e-7
