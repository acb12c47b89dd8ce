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
