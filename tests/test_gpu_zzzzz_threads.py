"""Threaded q-loop of lanczos_sqw / kpm_sqw (q_threads > 1: one context + model copy per host thread on the same GPU,
the structure of the reference's Threads.@threads loops, LanczosSqw.jl:65 / KPM_Sqw.jl:218).  Kept in its own file, run
last: it is the only test that calls the library from several host threads at once."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import sd  # noqa: E402


def test_threaded_q_loop_is_bit_identical():
    """The reference threads the q-loop of lanczos_sqw / kpm_sqw (LanczosSqw.jl:65, KPM_Sqw.jl:218); q_threads > 1
    does the same with one context per host thread on the same GPU and must reproduce the sequential loop exactly."""
    L = 12
    m = sd.XXZChain(L, nup=L // 2)
    _, psi0 = sd.groundstate(m, lanc_m=60, v0=np.random.default_rng(4).standard_normal(m.dim))
    q, w = sd.momenta(m), np.linspace(0.0, 4.0, 50)
    S1 = sd.lanczos_sqw(psi0, m, q, w, lanc_m=40, eta=0.1, q_batch=False)      # the sequential q-loop (the default batches the momenta)
    S4 = sd.lanczos_sqw(psi0, m, q, w, lanc_m=40, eta=0.1, q_threads=4)
    assert S1.shape == S4.shape == (L, 50) and np.array_equal(S1, S4)
    K1 = sd.kpm_sqw(psi0, m, q[:5], w, a=4.5, b=-1.0, kpm_m=64, q_batch=False)
    K3 = sd.kpm_sqw(psi0, m, q[:5], w, a=4.5, b=-1.0, kpm_m=64, q_threads=3)
    assert np.array_equal(K1, K3)
    S2 = sd.dynamical_structure_factor(m, psi0, q, w, method="lanczos", lanc_m=40, eta=0.1, q_threads=2)
    assert np.array_equal(S1, S2)
