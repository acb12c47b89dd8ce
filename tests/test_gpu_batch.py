"""The one-kernel-per-step recurrences of sd_batch.cu against the oracle and against the one-call-per-operation paths:
  * q-batched lanczos_sqw / kpm_sqw (LanczosSqw.jl:65-77, KPM_Sqw.jl:218-253 as one [state][q] multi-vector)
  * lanczos_groundstate's full reorthogonalisation as one cooperative kernel per step (Lanczos.jl:116-155)."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.oracle as orc  # noqa: E402
from conftest import sd  # noqa: E402


def _models(L, nup, boundary="open", **kw):
    return (sd.XXZChain(L, nup=nup, boundary=boundary, **kw), orc.XXZChain(L, nup=nup, boundary=boundary, **kw))


def _tridiag_batch(m, psi, q, lanc_m):
    from spindyn.core import _ptr
    dpsi = m.to_device(np.asarray(psi))
    n = len(q)
    alpha = np.zeros((n, lanc_m)); beta = np.zeros((n, lanc_m))
    meff = np.zeros(n, dtype=np.int32); nphi = np.zeros(n)
    qa = np.ascontiguousarray(q, dtype=np.float64)
    sd._lib.check(sd.lib().sd_lanczos_tridiag_szq_batch(m._h, dpsi._h, _ptr(qa), n, lanc_m, 1e-12, _ptr(alpha), _ptr(beta),
                                                      _ptr(meff), _ptr(nphi)))
    return alpha, beta, meff, nphi


@pytest.mark.parametrize("L,nup,boundary,nq", [(10, 5, "open", 10), (12, 6, "periodic", 7), (8, None, "open", 5), (16, 8, "open", 16),
                                               (14, 3, "open", 2), (12, 6, "open", 20)])
@pytest.mark.parametrize("cplx", [False, True])
def test_batched_tridiag_matches_oracle_per_momentum(L, nup, boundary, nq, cplx):
    """Every column of the multi-vector runs lanczos_tridiag of phi_q = Sz_q psi0: alpha / beta / norm within 1e-9 of the
    oracle's per-momentum call (lanc_m below the noise threshold, see test_gpu_sqw_tolerance.py); q = 0 in an Sz = 0
    sector gives norm(phi) = 0 -> m_eff = 0 (LanczosSqw.jl:69-72).  Block-layout, generic, full-basis and periodic models;
    column counts that pad (7 -> 8, 5 -> 6, 10 -> 12, 20 -> 24)."""
    m, om = _models(L, nup, boundary, Jxy=1.0, Jz=0.7, hz=0.0)
    rng = np.random.default_rng(L * 100 + nq)
    N = len(om)
    psi = rng.standard_normal(N) + (1j * rng.standard_normal(N) if cplx else 0.0)
    psi /= np.linalg.norm(psi)
    q = 2 * np.pi * np.arange(nq) / nq
    lanc_m = 8
    alpha, beta, meff, nphi = _tridiag_batch(m, psi, q, lanc_m)
    for c, qq in enumerate(q):
        phi = orc.Sz_q_vector(om, psi.astype(np.complex128), float(qq))
        n_ref = float(np.linalg.norm(phi))
        assert abs(nphi[c] - n_ref) <= 1e-12 * max(1.0, n_ref), (c, nphi[c], n_ref)
        if n_ref < 1e-13:
            assert meff[c] == 0 or nphi[c] < 1e-13
            continue
        a_ref, b_ref, _ = orc.lanczos_tridiag(orc.apply_H_, om, phi, lanc_m=lanc_m)
        k = int(meff[c])
        assert k == len(a_ref), (c, k, len(a_ref))
        assert np.allclose(alpha[c, :k], a_ref, rtol=1e-9, atol=1e-10), (c, alpha[c, :k] - a_ref)
        assert np.allclose(beta[c, :k - 1], b_ref, rtol=1e-9, atol=1e-10), (c, beta[c, :k - 1] - b_ref)


def test_batched_sqw_equals_the_q_loop_and_the_oracle():
    L, nup = 12, 6
    m, om = _models(L, nup)
    _, psi = orc.groundstate(om, lanc_m=60, v0=np.random.default_rng(3).standard_normal(len(om)))
    q = orc.momenta(om)
    w = np.linspace(0.0, 4.0, 64)
    S_ref = np.asarray(orc.lanczos_sqw(psi, om, q, w, lanc_m=10, eta=0.05))
    S_b = sd.lanczos_sqw(psi, m, q, w, lanc_m=10, eta=0.05, q_batch=True)
    S_l = sd.lanczos_sqw(psi, m, q, w, lanc_m=10, eta=0.05, q_batch=False)
    assert np.linalg.norm(S_b - S_ref) <= 1e-9 * np.linalg.norm(S_ref)
    assert np.linalg.norm(S_l - S_ref) <= 1e-9 * np.linalg.norm(S_ref)
    # Gauss broadening and a run-to-run identical result (deterministic reductions, test_Lanczos.jl:122-166)
    S_b2 = sd.lanczos_sqw(psi, m, q, w, lanc_m=10, eta=0.05, q_batch=True)
    assert np.array_equal(S_b, S_b2)


@pytest.mark.parametrize("L,nup,boundary,nq", [(10, 5, "open", 10), (12, 6, "periodic", 5), (16, 8, "open", 16)])
def test_batched_kpm_moments_match_oracle(L, nup, boundary, nq):
    """compute_chebyshev_moments (KPM_Sqw.jl:95-128) of phi_q / ||phi_q|| per column: 1e-11 like the per-momentum test."""
    from spindyn.core import _ptr
    m, om = _models(L, nup, boundary)
    rng = np.random.default_rng(17 + L)
    psi = rng.standard_normal(len(om)); psi /= np.linalg.norm(psi)
    q = 2 * np.pi * np.arange(nq) / nq
    M = 40
    a, b = (L / 4 + 0.5 + 0.45 * L) / (2 * 0.99) * 1.2, 0.0
    dpsi = m.to_device(psi)
    mu = np.zeros((nq, M)); nphi = np.zeros(nq); blown = ctypes.c_int()
    sd._lib.check(sd.lib().sd_kpm_moments_szq_batch(m._h, dpsi._h, _ptr(q.copy()), nq, M, a, b, _ptr(mu), _ptr(nphi), ctypes.byref(blown)))
    assert blown.value == 0
    for c, qq in enumerate(q):
        phi = orc.Sz_q_vector(om, psi.astype(np.complex128), float(qq))
        n_ref = float(np.linalg.norm(phi))
        assert abs(nphi[c] - n_ref) <= 1e-12 * max(1.0, n_ref)
        if n_ref < 1e-13:
            continue
        mu_ref = orc.compute_chebyshev_moments(orc.apply_H_, phi / n_ref, M, a, b, om)
        assert np.allclose(mu[c], mu_ref, rtol=0, atol=1e-11), (c, np.abs(mu[c] - mu_ref).max())


def test_batched_kpm_sqw_equals_the_q_loop():
    L, nup = 12, 6
    m, om = _models(L, nup)
    _, psi = orc.groundstate(om, lanc_m=60, v0=np.random.default_rng(4).standard_normal(len(om)))
    q = orc.momenta(om)
    w = np.linspace(0.0, 4.0, 50)
    a, b = 6.0, -0.5
    S_b = sd.kpm_sqw(psi, m, q, w, a=a, b=b, kpm_m=64, q_batch=True)
    S_l = sd.kpm_sqw(psi, m, q, w, a=a, b=b, kpm_m=64, q_batch=False)
    S_ref = np.asarray(orc.kpm_sqw(psi, om, q, w, a=a, b=b, kpm_m=64))
    assert np.linalg.norm(S_b - S_ref) <= 1e-9 * np.linalg.norm(S_ref)
    assert np.linalg.norm(S_l - S_ref) <= 1e-9 * np.linalg.norm(S_ref)


def test_batched_kpm_reports_blow_up_and_falls_back():
    """Rescaling bounds that are too tight: ||v_next|| grows past 1e3, where the reference renormalises (KPM_Sqw.jl:117-121).
    The batched kernel flags it and kpm_sqw falls back to the per-momentum path, which follows the reference."""
    from spindyn.core import _ptr
    L, nup = 10, 5
    m, om = _models(L, nup)
    psi = np.random.default_rng(8).standard_normal(len(om)); psi /= np.linalg.norm(psi)
    q = np.array([np.pi / 2, np.pi])
    dpsi = m.to_device(psi)
    mu = np.zeros((2, 60)); nphi = np.zeros(2); blown = ctypes.c_int()
    sd._lib.check(sd.lib().sd_kpm_moments_szq_batch(m._h, dpsi._h, _ptr(q.copy()), 2, 60, 1.0, 0.0, _ptr(mu), _ptr(nphi), ctypes.byref(blown)))
    assert blown.value == 1
    w = np.linspace(0.0, 1.0, 8)
    S = sd.kpm_sqw(psi, m, q, w, a=1.0, b=0.0, kpm_m=60)
    S_ref = np.asarray(orc.kpm_sqw(psi, om, q, w, a=1.0, b=0.0, kpm_m=60))
    assert np.allclose(S, S_ref, rtol=1e-6, atol=1e-9)


def test_batch_rejects_bad_arguments():
    from spindyn.core import _ptr
    m, _ = _models(8, 4)
    d = m.to_device(np.ones(m.dim))
    z = np.zeros(4)
    rc = sd.lib().sd_lanczos_tridiag_szq_batch(m._h, d._h, _ptr(z), 0, 5, 1e-12, _ptr(z), _ptr(z), _ptr(np.zeros(4, dtype=np.int32)), _ptr(z))
    assert rc != 0


# ------------------------------------------------------------------ fused reorthogonalisation

def _gs_tridiag(m, v0, lanc_m, orth_tol=1e-10):
    E0, psi, a, b = sd.lanczos_groundstate(sd.apply_H_, m, lanc_m=lanc_m, v0=v0, return_tridiag=True)
    return E0, psi, a, b


@pytest.mark.parametrize("L,nup,lanc_m", [(10, 5, 40), (12, 6, 100), (16, 8, 100), (8, 4, 70), (6, 3, 20)])
def test_fused_reorthogonalisation_matches_oracle_and_stepwise_path(L, nup, lanc_m, monkeypatch):
    """lanczos_groundstate with the cooperative one-kernel step (default) against the oracle (E0 1e-10, Ritz vector up
    to sign 1e-8, alpha / beta 1e-9 while they are well conditioned) and against the one-call-per-BLAS-operation path."""
    m, om = _models(L, nup)
    v0 = np.random.default_rng(L + lanc_m).standard_normal(len(om))
    E_ref, psi_ref = orc.lanczos_groundstate(orc.apply_H_, om, lanc_m=lanc_m, v0=v0)
    E_f, psi_f, a_f, b_f = _gs_tridiag(m, v0, lanc_m)
    monkeypatch.setenv("SD_REORTH_FUSED", "0")
    E_s, psi_s, a_s, b_s = _gs_tridiag(m, v0, lanc_m)
    monkeypatch.delenv("SD_REORTH_FUSED")
    assert abs(E_f - E_ref) < 1e-10 and abs(E_s - E_ref) < 1e-10, (E_f, E_s, E_ref)
    assert abs(np.linalg.norm(psi_f) - 1.0) < 1e-12
    sgn = np.sign(np.dot(psi_f, psi_ref))
    assert np.linalg.norm(sgn * psi_f - psi_ref) < 1e-7
    k = min(len(a_f), len(a_s), 12)
    assert np.allclose(a_f[:k], a_s[:k], rtol=1e-9, atol=1e-10) and np.allclose(b_f[:k - 1], b_s[:k - 1], rtol=1e-9, atol=1e-10)
    assert len(a_f) == len(a_s)
    # run-to-run identical (test_Lanczos.jl:122-166)
    E_f2, psi_f2, a_f2, _ = _gs_tridiag(m, v0, lanc_m)
    assert E_f2 == E_f and np.array_equal(psi_f2, psi_f) and np.array_equal(a_f2, a_f)


def test_fused_reorthogonalisation_launch_count():
    """config 1's ground state: one apply + one cooperative kernel per step instead of ~5 j BLAS-1 launches."""
    m, _ = _models(16, 8)
    v0 = np.random.default_rng(1).standard_normal(m.dim)
    l0 = m.ctx.launch_count()
    sd.groundstate(m, lanc_m=100, v0=v0)
    n = m.ctx.launch_count() - l0
    assert n < 400, n
