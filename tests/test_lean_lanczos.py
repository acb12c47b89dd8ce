"""The algorithm behind sd_lanczos_lean / lanczos_groundstate_lean (ground state on three work
vectors instead of the N x m basis of Lanczos.jl:104), restated in numpy on top of the oracle's apply_H! and gated
against the reference-faithful lanczos_groundstate: E0 within 1e-10, small residual, also far beyond the point where
the unorthogonalised recurrence produces copies of converged Ritz values."""
import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

import oracle.oracle as orc


def lean_groundstate(m, v0, lanc_m, tol=1e-12):
    """Mirror of sd_lanczos_lean's two passes (same loop structure, same index conventions)."""
    N = len(m)
    mm = min(lanc_m, N)

    def run(alpha=None, beta=None, y=None):
        first = y is None
        steps = mm if first else len(alpha)
        vj, vo, w = v0 / np.linalg.norm(v0), np.zeros(N), np.empty(N)
        a, b = [], []
        out = None if first else y[0] * vj
        for j in range(1, steps + 1):
            if not first and j == steps:
                break
            orc.apply_H_(w, vj, m)
            al = float(vj @ w) if first else alpha[j - 1]
            w -= al * vj
            if j > 1:
                w -= (b[-1] if first else beta[j - 2]) * vo
            if first:
                a.append(al)
                if j == steps:
                    break
                be = float(np.linalg.norm(w))
                if be < tol:
                    break
                b.append(be)
            else:
                be = beta[j - 1]
            vo, vj, w = vj, w / be, np.empty(N)
            if not first:
                out += y[j] * vj
        return (np.array(a), np.array(b)) if first else out

    a, b = run()
    theta, Q = eigh_tridiagonal(a, b[:len(a) - 1]) if len(a) > 1 else (a, np.ones((1, 1)))
    psi = run(a, b, Q[:, 0])
    return float(theta[0]), psi / np.linalg.norm(psi), len(a)


@pytest.mark.parametrize("L,nup,lanc_m", [(8, 4, 30), (10, 5, 60), (12, 6, 80), (12, 4, 200), (14, 7, 120), (6, 3, 100)])
def test_two_pass_lean_ground_state_equals_the_faithful_one(L, nup, lanc_m):
    m = orc.XXZChain(L, Jxy=1.0, Jz=1.0, nup=nup)
    v0 = np.random.default_rng(L + lanc_m).standard_normal(len(m))
    E, psi, k = lean_groundstate(m, v0, lanc_m)
    Eref, pref = orc.lanczos_groundstate(orc.apply_H_, m, lanc_m=lanc_m, v0=v0)
    assert abs(E - Eref) < 1e-10
    h = np.empty_like(psi)
    orc.apply_H_(h, psi, m)
    assert abs(psi @ h - E) < 1e-10 and np.linalg.norm(h - E * psi) < 1e-6
    assert min(np.linalg.norm(psi - pref), np.linalg.norm(psi + pref)) < 1e-6      # up to LAPACK's sign


def test_lean_breakdown_on_an_invariant_start_vector():
    """beta < tol ends pass 1 early (Lanczos.jl:136-139 semantics); pass 2 then uses m_eff vectors only."""
    m = orc.XXZChain(8, nup=4)
    _, gs = orc.lanczos_groundstate(orc.apply_H_, m, lanc_m=70, v0=np.random.default_rng(0).standard_normal(len(m)))
    E, psi, k = lean_groundstate(m, gs, 40, tol=1e-8)
    assert k < 40 and abs(E - (-3.374932598687896)) < 1e-10 and abs(abs(psi @ gs) - 1) < 1e-10


def deferred_recurrence(m, v0, steps, hsign=1.0):
    """Mirror of sd_lanczos_engine (pass 1): unnormalised vectors u_1 = v0, u_{j+1} = w_j, v_j = u_j / beta_{j-1};
    apply w = (hsign / beta_{j-1}) H u_j with d_j = <u_j, w>; update w -= (alpha_j / beta_{j-1}) u_j +
    (beta_{j-1} / beta_{j-2}) u_{j-1}; alpha_j = d_j / beta_{j-1}, beta_j = ||w||.  No vector is ever divided."""
    N = len(m)
    u, uo, w = v0.copy(), np.zeros(N), np.empty(N)
    n = [float(v0 @ v0)]
    alpha, beta = [], []
    for j in range(1, steps + 1):
        b1 = np.sqrt(n[j - 1])
        orc.apply_H_(w, u, m)
        w *= hsign / b1
        d = float(u @ w)
        al = d / b1
        alpha.append(al)
        if j == steps:
            break
        w -= (al / b1) * u
        if j >= 2:
            w -= (b1 / np.sqrt(n[j - 2])) * uo
        n.append(float(w @ w))
        beta.append(np.sqrt(n[j]))
        uo, u, w = u, w, uo
    return np.array(alpha), np.array(beta)


@pytest.mark.parametrize("L,nup,steps", [(8, 4, 20), (10, 5, 40), (12, 6, 60)])
def test_deferred_normalisation_gives_the_same_tridiagonal_matrix(L, nup, steps):
    """The fused recurrence of sd_lanczos_engine against the textbook one (oracle lanczos_tridiag, Lanczos.jl:196-246)."""
    m = orc.XXZChain(L, Jxy=1.0, Jz=1.0, nup=nup)
    rng = np.random.default_rng(L)
    v0 = 3.7 * rng.standard_normal(len(m))                                       # not normalised on purpose
    a, b = deferred_recurrence(m, v0, steps)
    ar, br, nv = orc.lanczos_tridiag(orc.apply_H_, m, v0.astype(np.complex128), lanc_m=steps)
    assert abs(nv - np.linalg.norm(v0)) < 1e-12
    k = min(len(a), len(ar), 15)              # before rounding noise separates two unorthogonalised recurrences
    assert np.allclose(a[:k], ar[:k], atol=1e-9) and np.allclose(b[:k - 1], br[:k - 1], atol=1e-9)
    ev = eigh_tridiagonal(a, b, eigvals_only=True)
    evr = eigh_tridiagonal(np.asarray(ar), np.asarray(br), eigvals_only=True)
    assert abs(ev[0] - evr[0]) < 1e-10 and abs(ev[-1] - evr[-1]) < 1e-10
    an, bn = deferred_recurrence(m, v0, steps, hsign=-1.0)                       # -H (estimate_energy_bounds, Lanczos.jl:261-265)
    assert abs(eigh_tridiagonal(an, bn, eigvals_only=True)[0] + ev[-1]) < 1e-10
