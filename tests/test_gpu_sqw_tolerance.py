"""Config 1 of BASELINE.json (XXZChain L=16 nup=8, lanczos_sqw over the 16 momenta, 100 frequencies, lanc_m=100,
eta=0.05; LanczosSqw.jl:49-80) through the GPU path against the oracle, with a tolerance that is MEASURED, not chosen.

Plain (unreorthogonalised) Lanczos amplifies rounding noise once Ritz values converge, so two correct implementations
that differ in the last bit of any intermediate (FMA contraction, summation order, deferred normalisation) do not agree
to 1e-9 in S(q,w) at lanc_m = 100.  The yardstick is the oracle's own sensitivity: the same oracle call with psi0
perturbed by one unit of relative rounding (1e-16 noise) moves S(q,w) by d_ref.  The GPU result must lie within that
envelope (a small multiple of the largest d_ref over several noise seeds); what IS well conditioned -- E0 and the first
Lanczos coefficients of every momentum's recurrence -- must agree tightly."""
import numpy as np
import pytest

import spindyn as sd
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_config1_sqw_within_the_oracles_own_rounding_envelope():
    L, nup, lanc_m, eta = 16, 8, 100, 0.05
    w = np.linspace(0.0, 5.0, 100)
    om = orc.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=nup)
    m = sd.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=nup)
    v0 = np.random.default_rng(5).standard_normal(len(om))
    E0, psi = orc.groundstate(om, lanc_m=100, v0=v0)
    q = orc.momenta(om)
    S_ref = np.asarray(orc.lanczos_sqw(psi, om, q, w, lanc_m=lanc_m, eta=eta))
    d_ref = 0.0
    for seed in range(3):
        noise = 1.0 + 1e-16 * np.random.default_rng(100 + seed).standard_normal(len(psi))
        S_p = np.asarray(orc.lanczos_sqw(psi * noise, om, q, w, lanc_m=lanc_m, eta=eta))
        d_ref = max(d_ref, float(np.linalg.norm(S_p - S_ref) / np.linalg.norm(S_ref)))
    S_gpu = np.asarray(sd.lanczos_sqw(psi, m, q, w, lanc_m=lanc_m, eta=eta))
    d_gpu = float(np.linalg.norm(S_gpu - S_ref) / np.linalg.norm(S_ref))
    tol = max(20.0 * d_ref, 1e-9)
    print(f"config 1 S(q,w): oracle self-sensitivity d_ref = {d_ref:.2e}, GPU vs oracle = {d_gpu:.2e}, tolerance = {tol:.2e}")
    assert d_gpu <= tol, f"GPU vs oracle {d_gpu:.3e} > 20 x oracle self-sensitivity {d_ref:.3e}"
    # what is well conditioned must agree tightly: E0 of the GPU ground state, and the head of the tridiagonal matrix
    Eg, _ = sd.groundstate(m, lanc_m=100, v0=v0)
    assert abs(Eg - E0) < 1e-10, (Eg, E0)
    for qq in q[1:4]:
        phi = orc.Sz_q_vector(om, psi, float(qq))
        a_ref, b_ref, n_ref = orc.lanczos_tridiag(orc.apply_H_, om, phi, lanc_m=12)
        a_gpu, b_gpu, n_gpu = sd.lanczos_tridiag(sd.apply_H_, m, phi, lanc_m=12)
        assert abs(n_gpu - n_ref) < 1e-12 * max(1.0, n_ref)
        assert np.allclose(a_gpu, a_ref, rtol=1e-9, atol=1e-10) and np.allclose(b_gpu, b_ref, rtol=1e-9, atol=1e-10), (a_gpu - a_ref, b_gpu - b_ref)


@pytest.mark.parametrize("lanc_m", [10])
def test_sqw_tight_parity_below_the_noise_threshold(lanc_m):
    """Before Ritz values converge the recurrence is well conditioned: 1e-9 relative, the north-star figure.  The oracle's
    own sensitivity to 1e-16 input noise for this very setup (ground-state-derived start vectors converge fast) is
    4e-14 / 9e-14 / 2e-12 / 4e-10 / 3e-8 at lanc_m = 8 / 10 / 12 / 16 / 20 and ~5e-5 at 100, so the tight bar is asserted
    at lanc_m = 10 and the measured envelope above takes over beyond."""
    L, nup = 16, 8
    w = np.linspace(0.0, 5.0, 100)
    om = orc.XXZChain(L, nup=nup)
    m = sd.XXZChain(L, nup=nup)
    _, psi = orc.groundstate(om, lanc_m=100, v0=np.random.default_rng(6).standard_normal(len(om)))
    q = orc.momenta(om)[:6]
    S_ref = np.asarray(orc.lanczos_sqw(psi, om, q, w, lanc_m=lanc_m, eta=0.05))
    S_gpu = np.asarray(sd.lanczos_sqw(psi, m, q, w, lanc_m=lanc_m, eta=0.05))
    assert np.linalg.norm(S_gpu - S_ref) <= 1e-9 * np.linalg.norm(S_ref)
