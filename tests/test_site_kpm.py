"""Site-resolved KPM (TimeEvolution/KPM.jl:74-206): the oracle restatement against exact
diagonalisation, and (GPU) the device drivers against the oracle."""
import numpy as np
import pytest

import dense_ref
import oracle.oracle as orc


def exact_cross_spectrum(H, chi, phi):
    """mu_k = chi^T T_k(H~) phi needs nothing but the eigen-decomposition."""
    ev, U = np.linalg.eigh(H)
    return ev, (U.T @ chi) * (U.T @ phi)


@pytest.mark.parametrize("L,nup,i,j", [(6, 3, 2, 2), (8, 4, 3, 5)])
def test_oracle_cross_moments_match_exact_diagonalisation(L, nup, i, j):
    m = orc.XXZChain(L, Jxy=1.0, Jz=0.7, nup=nup)
    H = dense_ref.dense_H(L, nup, *dense_ref.xxz_lists(L, Jxy=1.0, Jz=0.7))
    ev = np.linalg.eigvalsh(H)
    a, b = (ev[-1] - ev[0]) / (2 * 0.95), (ev[-1] + ev[0]) / 2
    psi = np.linalg.eigh(H)[1][:, 0]
    chi = orc.site_sz_operator(i)(psi, m)
    phi = orc.site_sz_operator(j)(psi, m)
    n = 40
    mu = orc.compute_cross_chebyshev_moments(chi, phi, n, a, b, orc.apply_H_, m)
    e, wts = exact_cross_spectrum(H, chi.real, phi.real)
    x = (e - b) / a
    want = np.array([np.sum(wts * np.cos(k * np.arccos(x))) for k in range(n)])
    assert np.allclose(mu, want, atol=1e-12)
    S = orc.kpm_dynamical_correlation(psi, orc.site_sz_operator(i), orc.site_sz_operator(j), np.linspace(ev[0], ev[-1], 50),
                                      orc.apply_H_, m, n=n, a=a, b=b)
    assert S.shape == (50,) and np.all(S >= 0.0) and np.isfinite(S).all()


def test_jackson_kernel_and_series_evaluation():
    g = orc.get_jackson_kernel(10)
    assert abs(g[0] - 1.0) < 1e-15 and np.all(np.diff(g) < 0) and g[-1] > 0
    assert orc.evaluate_chebyshev_series(np.array([1.0, 0.0, 0.0]), 1.2, 2.0) == 0.0
    assert abs(orc.evaluate_chebyshev_series(np.array([1.0, 0.5]), 0.3, 2.0) - (1 + 0.15) / (np.pi * np.sqrt(1 - 0.09))) < 1e-15


@pytest.mark.gpu
@pytest.mark.parametrize("L,nup,bc", [(10, 5, "open"), (16, 8, "open"), (12, None, "open")])
def test_gpu_site_resolved_kpm_matches_oracle(L, nup, bc):
    import spindyn as sd
    m = sd.XXZChain(L, Jxy=1.0, Jz=0.8, hz=0.1, nup=nup, boundary=bc)
    om = orc.XXZChain(L, Jxy=1.0, Jz=0.8, hz=0.1, nup=nup, boundary=bc)
    rng = np.random.default_rng(L)
    psi = rng.standard_normal(len(om))
    psi /= np.linalg.norm(psi)
    a, b = L / 3.0, -0.1
    for (i, j) in [(1, 1), (2, L - 1), (L, 3)]:
        chi, phi = orc.site_sz_operator(i)(psi, om), orc.site_sz_operator(j)(psi, om)
        got_phi = sd.site_sz_operator(j)(psi, m)
        assert np.allclose(got_phi, phi, atol=1e-15)
        mu_ref = orc.compute_cross_chebyshev_moments(chi, phi, 48, a, b, orc.apply_H_, om)
        mu = sd.compute_cross_chebyshev_moments(chi, phi, 48, a, b, sd.apply_H_, m)
        assert np.allclose(mu, mu_ref, atol=1e-11)
    w = np.linspace(-2.0, 2.0, 64)
    S_ref = orc.kpm_dynamical_correlation(psi, orc.site_sz_operator(2), orc.site_sz_operator(3), w, orc.apply_H_, om, n=64, a=a, b=b)
    S = sd.kpm_dynamical_correlation(psi, sd.site_sz_operator(2), sd.site_sz_operator(3), w, sd.apply_H_, m, n=64, a=a, b=b)
    assert np.allclose(S, S_ref, rtol=1e-9, atol=1e-11)
    if L <= 10:
        C = sd.kpm_correlation_matrix(psi, w[:8], sd.apply_H_, m, n=16)
        assert C.shape == (L, L, 8) and np.all(C >= 0)
        assert sd.Sqw(C, 0.3, np.arange(L, dtype=float)).shape == (8,)
