"""GPU mirrors of the reference's solver tests (test_Lanczos.jl, test_KPM.jl,
test_PublicAPI.jl, test_InitialStates.jl) run through the C ABI, plus
comparisons of every recurrence with the CPU oracle on identical inputs.
Tolerances follow the north star: E0 1e-10, psi(t) and S(q,w) 1e-9 relative,
and the reference tests' own atol where they are tighter."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.oracle as orc  # noqa: E402
from conftest import sd  # noqa: E402
from dense_ref import dense_H, xxz_lists  # noqa: E402


def dense_from_apply(m):
    """test_Lanczos.jl:36-44: H built column by column from apply_H! itself."""
    N = m.dim
    H = np.zeros((N, N))
    out = np.zeros(N)
    for j in range(N):
        e = np.zeros(N)
        e[j] = 1.0
        sd.apply_H_(out, e, m)
        H[:, j] = out
    return H


# ------------------------------------------------------------ test_Lanczos.jl

def test_lanczos_tridiag_complex_alpha():
    """test_Lanczos.jl:6-26."""
    m = sd.XXZChain(2, Jxy=1.0, Jz=1.0, nup=1)
    v = np.array([1.0, 1j]) / np.sqrt(2)
    Hv = np.empty_like(v)
    sd.apply_H_(Hv, v, m)
    alpha, beta, normv = sd.lanczos_tridiag(sd.apply_H_, m, v, lanc_m=2)
    assert abs(alpha[0] - np.vdot(v, Hv).real) < 1e-12
    assert abs(normv - 1.0) < 1e-12


def test_groundstate_vs_exact_diagonalisation():
    """test_Lanczos.jl:29-54 (L=6, lanc_m=N) + independent Kronecker H."""
    m = sd.XXZChain(6, Jxy=1.0, Jz=1.0, nup=3)
    H = dense_from_apply(m)
    assert np.allclose(H, dense_H(6, 3, *xxz_lists(6)), atol=1e-14)
    E_exact = np.linalg.eigvalsh(H).min()
    E0, psi0 = sd.groundstate(m, lanc_m=m.dim)
    assert abs(E0 - E_exact) < 1e-12
    assert abs(np.linalg.norm(psi0) - 1.0) < 1e-12
    assert np.linalg.norm(H @ psi0 - E0 * psi0) < 1e-10


def test_lanczos_dimension_capped():
    """test_Lanczos.jl:57-119."""
    m = sd.XXZChain(4, nup=2)
    E0, psi0 = sd.groundstate(m, lanc_m=100)
    assert len(psi0) == m.dim and abs(np.linalg.norm(psi0) - 1) < 1e-12
    Hpsi = np.empty_like(psi0)
    sd.apply_H_(Hpsi, psi0, m)
    assert np.linalg.norm(Hpsi - E0 * psi0) < 1e-10
    exact = np.linalg.eigvalsh(dense_from_apply(m))
    Emin, Emax = sd.lanczos_extremal(sd.apply_H_, m, lanc_m=100)
    assert abs(Emin - exact[0]) < 1e-12 and abs(Emax - exact[-1]) < 1e-12
    rng = np.random.default_rng(0)
    v = sd.randn_complex(rng, m.dim)
    a, b, _ = sd.lanczos_tridiag(sd.apply_H_, m, v / np.linalg.norm(v), lanc_m=100)
    assert len(a) <= m.dim and len(b) == len(a) - 1


def test_reproducibility_with_explicit_rng():
    """test_Lanczos.jl:122-166: same seed -> identical results (fixed-order reductions)."""
    m = sd.XXZChain(6, Jxy=1.0, Jz=1.0, nup=3)
    E1, p1 = sd.groundstate(m, lanc_m=10, rng=np.random.default_rng(1234))
    E2, p2 = sd.groundstate(m, lanc_m=10, rng=np.random.default_rng(1234))
    assert abs(E1 - E2) <= 1e-14 and np.max(np.abs(p1 - p2)) <= 1e-14
    b1 = sd.lanczos_extremal(sd.apply_H_, m, lanc_m=10, rng=np.random.default_rng(42))
    b2 = sd.lanczos_extremal(sd.apply_H_, m, lanc_m=10, rng=np.random.default_rng(42))
    assert np.allclose(b1, b2, atol=1e-14, rtol=0)


def test_zero_start_vector_error():
    m = sd.XXZChain(4, nup=2)
    with pytest.raises(RuntimeError, match="zero norm"):
        sd.lanczos_tridiag(sd.apply_H_, m, np.zeros(m.dim, dtype=np.complex128))


# ------------------------------------------------------------ vs the oracle

@pytest.mark.parametrize("L,nup,boundary", [(12, 6, "open"), (16, 8, "open"), (10, 5, "periodic"), (10, None, "open"), (16, 8, "periodic")])
def test_recurrences_match_oracle(L, nup, boundary):
    m = sd.XXZChain(L, Jxy=1.0, Jz=0.8, hz=0.05, nup=nup, boundary=boundary)
    om = orc.XXZChain(L, Jxy=1.0, Jz=0.8, hz=0.05, nup=nup, boundary=boundary)
    rng = np.random.default_rng(L)
    N = m.dim
    vc = sd.randn_complex(rng, N)
    vr = rng.standard_normal(N)
    # lanczos_extremal / estimate bounds on H and -H
    assert np.allclose(sd.lanczos_extremal(sd.apply_H_, m, lanc_m=40, v0=vc),
                       orc.lanczos_extremal(orc.apply_H_, om, lanc_m=40, v0=vc), atol=1e-10, rtol=0)
    # lanczos_tridiag coefficients
    a1, b1, n1 = sd.lanczos_tridiag(sd.apply_H_, m, vc, lanc_m=25)
    a2, b2, n2 = orc.lanczos_tridiag(orc.apply_H_, om, vc, lanc_m=25)
    assert np.allclose(a1, a2, atol=1e-9) and np.allclose(b1, b2, atol=1e-9) and abs(n1 - n2) < 1e-12
    # ground state with full reorthogonalisation
    E1, p1, al1, be1 = sd.lanczos_groundstate(sd.apply_H_, m, lanc_m=60, v0=vr, return_tridiag=True)
    E2, p2, al2, be2 = orc.lanczos_groundstate(orc.apply_H_, om, lanc_m=60, v0=vr, return_tridiag=True)
    assert abs(E1 - E2) < 1e-10
    assert min(np.linalg.norm(p1 - p2), np.linalg.norm(p1 + p2)) < 1e-7        # Ritz vector up to LAPACK's sign
    Hp = np.empty_like(p1)
    sd.apply_H_(Hp, p1, m)
    assert abs(np.vdot(p1, Hp) - E1) < 1e-10
    # KPM moments
    a, b = 1.05 * (L / 4 + 1), 0.1
    phi = vc / np.linalg.norm(vc)
    mu1 = sd.compute_chebyshev_moments(sd.apply_H_, phi, 64, a, b, m)
    mu2 = orc.compute_chebyshev_moments(orc.apply_H_, phi, 64, a, b, om)
    assert np.allclose(mu1, mu2, atol=1e-11, rtol=0)
    # Krylov and Chebyshev time evolution
    k1 = sd.krylov_time_evolve(vc, 0.37, sd.apply_H_, m, kry_m=20)
    k2 = orc.krylov_time_evolve(vc, 0.37, orc.apply_H_, om, kry_m=20)
    assert np.linalg.norm(k1 - k2) < 1e-9
    kr1 = sd.krylov_time_evolve(vr, 0.2, sd.apply_H_, m, kry_m=12)
    kr2 = orc.krylov_time_evolve(vr, 0.2, orc.apply_H_, om, kry_m=12)
    assert np.linalg.norm(kr1 - kr2) < 1e-9
    bounds = (-(L / 4 + 1.0), L / 4 + 1.0)
    c1 = sd.chebyshev_time_evolve(vc, 0.25, sd.apply_H_, m, cheb_n=40, Ebounds=bounds)
    c2 = orc.chebyshev_time_evolve(vc, 0.25, orc.apply_H_, om, cheb_n=40, Ebounds=bounds)
    assert np.linalg.norm(c1 - c2) < 1e-9 * np.linalg.norm(c2)


def test_sqw_matches_oracle():
    L, nup = 10, 5
    m, om = sd.XXZChain(L, nup=nup), orc.XXZChain(L, nup=nup)
    rng = np.random.default_rng(8)
    E2, psi0 = orc.lanczos_groundstate(orc.apply_H_, om, lanc_m=80, v0=rng.standard_normal(m.dim))
    q = sd.momenta(m)
    w = np.linspace(0.0, 4.0, 60)
    S1 = sd.dynamical_structure_factor(m, psi0, q, w, method="lanczos", lanc_m=20, eta=0.05)
    S2 = orc.dynamical_structure_factor(om, psi0, q, w, method="lanczos", lanc_m=20, eta=0.05)
    assert S1.shape == (L, 60)
    assert np.max(np.abs(S1 - S2)) < 1e-9 * max(1.0, np.max(np.abs(S2)))
    a, b = sd._rescaling_from_bounds(-(L / 4 + 1.0), L / 4 + 1.0)
    K1 = sd.dynamical_structure_factor(m, psi0, q[:4], w, method="kpm", kpm_m=80, a=a, b=b)
    K2 = orc.dynamical_structure_factor(om, psi0, q[:4], w, method="kpm", kpm_m=80, a=a, b=b)
    assert np.max(np.abs(K1 - K2)) < 1e-9 * max(1.0, np.max(np.abs(K2)))


# ------------------------------------------------------------ test_PublicAPI.jl / test_KPM.jl

def test_public_groundstate_L2():
    """test_PublicAPI.jl:40-53."""
    m = sd.XXZChain(2, Jxy=1.0, Jz=1.0, nup=1)
    E0, psi0 = sd.groundstate(m, lanc_m=2)
    assert abs(E0 + 0.75) < 1e-12 and abs(np.linalg.norm(psi0) - 1) < 1e-12
    Hp = np.empty_like(psi0)
    sd.apply_H_(Hp, psi0, m)
    assert np.linalg.norm(Hp - E0 * psi0) < 1e-10
    with pytest.raises(ValueError):
        sd.groundstate(m, method="unknown")
    assert np.allclose(sd.momenta(sd.XXZChain(6, nup=3)), 2 * np.pi * np.arange(6) / 6)


def test_public_time_evolve():
    """test_PublicAPI.jl:56-134."""
    from scipy.linalg import expm
    m = sd.XXZChain(2, Jxy=1.0, Jz=1.0, nup=1)
    H = np.array([[-0.25, 0.5], [0.5, -0.25]])
    psi0 = np.array([1.0 + 0j, 0.0])
    t = 0.3
    exact = expm(-1j * t * H) @ psi0
    pk = sd.time_evolve(m, psi0, t, method="krylov", kry_m=2)
    assert np.allclose(pk, exact, atol=1e-10) and abs(np.linalg.norm(pk) - 1) < 1e-12
    assert np.allclose(sd.time_evolve(m, psi0, 0.0, method="krylov", kry_m=2), psi0, atol=1e-12)
    with pytest.raises(ValueError):
        sd.time_evolve(m, psi0, t, method="unknown")
    pc = sd.time_evolve(m, psi0, t, method="chebyshev", cheb_n=30, Ebounds=(-0.75, 0.25))
    assert np.allclose(pc, exact, atol=1e-8) and abs(np.linalg.norm(pc) - 1) < 1e-8
    pa = sd.time_evolve(m, psi0, 0.1, method="chebyshev", cheb_n=20)
    assert abs(np.linalg.norm(pa) - 1) < 1e-6
    with pytest.raises(TypeError):
        sd.time_evolve(m, np.array([1.0, 0.0]), t, method="chebyshev", Ebounds=(-1, 1))   # InexactError


def test_public_dynamical_structure_factor():
    """test_PublicAPI.jl:154-203."""
    m = sd.XXZChain(4, Jxy=1.0, Jz=1.0, nup=2)
    _, psi0 = sd.groundstate(m, lanc_m=6)
    q = sd.momenta(m)
    w = np.linspace(0.0, 3.0, 40)
    S = sd.dynamical_structure_factor(m, psi0, q, w, method="lanczos", lanc_m=6, eta=0.05)
    assert S.shape == (4, 40) and np.all(np.isfinite(S)) and np.all(S >= -1e-12)
    with pytest.raises(ValueError):
        sd.dynamical_structure_factor(m, psi0, q, w, method="unknown")
    S = sd.dynamical_structure_factor(m, psi0, q, np.linspace(-2, 2, 40), method="kpm", kpm_m=40)
    assert S.shape == (4, 40) and np.all(np.isfinite(S))


def test_kpm_scale_and_sum_rule():
    """test_KPM.jl:44-91."""
    m = sd.XXZChain(6, Jxy=1.0, Jz=1.0, nup=3)
    _, psi0 = sd.groundstate(m, lanc_m=20)
    w = np.linspace(0.0, 5.0, 300)
    S = sd.dynamical_structure_factor(m, psi0, [np.pi], w, method="kpm", kpm_m=100)
    assert np.all(np.isfinite(S)) and np.all(S >= 0) and S[:, -11:].max() < S.max()
    w = np.arange(0.0, 5.0 + 1e-9, 0.01)
    phi = sd.Sz_q_vector(m, psi0, np.pi)
    exact = np.linalg.norm(phi) ** 2
    S = sd.dynamical_structure_factor(m, psi0, [np.pi], w, method="kpm", kpm_m=120, kernel="jackson")
    assert abs(S[0].sum() * 0.01 - exact) < 5e-3 * exact
    a, b = sd.get_rescaling_params(sd.apply_H_, m)                  # test_KPM.jl:4-29
    ev = np.linalg.eigvalsh(dense_from_apply(m))
    assert -1 < (ev[0] - b) / a < 0 < (ev[-1] - b) / a < 1


def test_open_heisenberg_ground_energies():
    """independent E0 of the open Heisenberg chain, Sz=0."""
    for L, E in [(8, -3.374932598687896), (12, -5.1420906328405), (16, -6.9117371455751)]:
        m = sd.XXZChain(L, nup=L // 2)
        E0, _ = sd.groundstate(m, lanc_m=min(m.dim, 90), rng=np.random.default_rng(L))
        assert abs(E0 - E) < 1e-10, (L, E0)


# ------------------------------------------------------------ test_InitialStates.jl (boundary: sd_rank)

def test_initial_states():
    full, sec = sd.XXZChain(4), sd.XXZChain(4, nup=2)
    n = sd.neel_state(full)
    assert n.dtype == np.float64 and n.sum() == 1.0 and n[0b0101] == 1.0          # :32-39
    assert sd.polarized_state(full)[15] == 1.0 and sd.polarized_state(full, up=False)[0] == 1.0   # :60-63
    assert sd.polarized_state_with_flips(full, [2, 4])[0b0101] == 1.0                 # :89-90
    assert sd.neel_state(sec).sum() == 1.0 and sd.domain_wall_state(sec)[0] == 1.0
    with pytest.raises(ValueError):
        sd.polarized_state(sec)                                                        # :72-73
    with pytest.raises(ValueError):
        sd.polarized_state_with_flips(sec, [1])
    with pytest.raises(ValueError):
        sd.polarized_state_with_flips(full, [5])
    om = orc.XXZChain(10, nup=5)
    m = sd.XXZChain(10, nup=5)
    assert np.array_equal(sd.neel_state(m), orc.neel_state(om))
    assert np.array_equal(sd.neel_state(m, device=True).to_host(), orc.neel_state(om))


def test_config2_krylov_from_neel_device_resident():
    """BASELINE config 2 at a size the oracle finishes in seconds (L=20 instead of 24)."""
    L = 20
    m, om = sd.XXZChain(L, nup=L // 2), orc.XXZChain(L, nup=L // 2)
    psi0 = sd.neel_state(m, device=True).astype(np.complex128)
    pt = sd.krylov_time_evolve(psi0, 0.5, sd.apply_H_, m, kry_m=30)
    ref = orc.krylov_time_evolve(orc.neel_state(om).astype(np.complex128), 0.5, orc.apply_H_, om, kry_m=30)
    assert np.linalg.norm(pt.to_host() - ref) < 1e-9


def test_inplace_krylov_and_workspaces():
    """krylov_time_evolve! / KrylovWorkspace (Krylov.jl:25-118) and ChebyshevWorkspace (Chebyshev.jl:19-36,83-87):
    same psi(t) as the out-of-place entry points, the reference's size assertions, zero-norm early return."""
    L = 10
    m, om = sd.XXZChain(L, nup=5), orc.XXZChain(L, nup=5)
    rng = np.random.default_rng(8)
    psi0 = sd.randn_complex(rng, m.dim)
    psi0 /= np.linalg.norm(psi0)
    ref = orc.krylov_time_evolve(psi0, 0.3, orc.apply_H_, om, kry_m=20)
    ws = sd.KrylovWorkspace(m.dim, 20)
    out = np.zeros(m.dim, dtype=np.complex128)
    assert sd.krylov_time_evolve_(out, psi0, 0.3, sd.apply_H_, m, ws, kry_m=20) is None
    assert np.linalg.norm(out - ref) < 1e-9
    with pytest.raises(ValueError):
        sd.krylov_time_evolve_(out, psi0, 0.3, sd.apply_H_, m, sd.KrylovWorkspace(m.dim, 5), kry_m=20)
    with pytest.raises(ValueError):
        sd.krylov_time_evolve_(np.zeros(3, dtype=np.complex128), psi0, 0.3, sd.apply_H_, m, ws, kry_m=20)
    z = np.zeros(m.dim, dtype=np.complex128)
    assert sd.krylov_time_evolve_(out, z, 0.3, sd.apply_H_, m, ws, kry_m=20) is out and not np.any(out)
    cw = sd.ChebyshevWorkspace(psi0)
    a = sd.chebyshev_time_evolve(psi0, 0.2, sd.apply_H_, m, cheb_n=40, Ebounds=(-5.0, 3.0), workspace=cw)
    b = sd.chebyshev_time_evolve(psi0, 0.2, sd.apply_H_, m, cheb_n=40, Ebounds=(-5.0, 3.0))
    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        sd.chebyshev_time_evolve(psi0, 0.2, sd.apply_H_, m, cheb_n=40, Ebounds=(-5.0, 3.0), workspace=sd.ChebyshevWorkspace(3))



def test_kpm_moments_renormalisation_path_matches_oracle():
    """KPM_Sqw.jl:117-121: with rescaling bounds that are too tight the recurrence grows and the reference renormalises
    v_next whenever its norm exceeds 1e3.  The device loop runs 32 moments speculatively and falls back to the
    step-by-step path from a checkpoint when a norm in the block crosses the threshold: same moments as the oracle."""
    L, nup = 12, 6
    m = sd.XXZChain(L, nup=nup)
    om = orc.XXZChain(L, nup=nup)
    rng = np.random.default_rng(11)
    phi = rng.standard_normal(m.dim) + 1j * rng.standard_normal(m.dim)
    phi /= np.linalg.norm(phi)
    a, b = 1.2, -0.3                                    # spectrum width ~ 8: |H~| > 1, the Chebyshev recurrence diverges
    M = 80
    mu_ref = np.asarray(orc.compute_chebyshev_moments(orc.apply_H_, phi.copy(), M, a, b, om))
    mu = np.asarray(sd.compute_chebyshev_moments(sd.apply_H_, phi, M, a, b, m))
    assert np.max(np.abs(mu_ref)) > 1e3                 # the path was really taken
    assert np.allclose(mu, mu_ref, rtol=1e-9, atol=1e-9)
