"""Pins the CPU oracle (oracle/) against the reference's own known answers and
against an independent Kronecker-product construction (tests/dense_ref.py).
CPU only; no CUDA involved."""
import itertools

import numpy as np
import pytest
from scipy.linalg import expm

import dense_ref
from oracle import oracle as orc

RNG = np.random.default_rng(20261018)


# ------------------------------------------------------------------ basis

def test_basis_validation_and_edges():
    """test_Basis.jl:4-19."""
    for bad in (0, -1):
        with pytest.raises(ValueError):
            orc.build_full_basis(bad)
    for bad in (-1, 5):
        with pytest.raises(ValueError):
            orc.build_sector_basis(4, bad)
    st, idx = orc.build_sector_basis(4, 0)
    assert len(st) == 1 and st[0] == 0 and idx[0] == 1
    st, idx = orc.build_sector_basis(4, 4)
    assert len(st) == 1 and bin(int(st[0])).count("1") == 4


def test_sector_order_hand_derived():
    """L=4,k=2 -> [3,5,9,6,10,12]; L=6,k=3 starts [7,11,19,35,13,21]."""
    st, _ = orc.build_sector_basis(4, 2)
    assert st.tolist() == [3, 5, 9, 6, 10, 12]
    st, _ = orc.build_sector_basis(6, 3)
    assert st[:6].tolist() == [7, 11, 19, 35, 13, 21]


@pytest.mark.parametrize("L", [1, 2, 3, 5, 8, 11, 14])
def test_c_basis_matches_itertools_and_closed_form(L):
    for nup in range(L + 1):
        ref = np.array(dense_ref.sector_states(L, nup), dtype=np.uint64)
        m = orc.build_model(L, nup=nup)
        assert np.array_equal(m.states, ref)
        # closed-form rank == position in the enumeration
        sel = ref if len(ref) < 400 else ref[RNG.integers(0, len(ref), 400)]
        for s in sel:
            r = orc.lib().orc_rank_closed_form(L, nup, int(s))
            assert ref[r] == s
        # get(idxmap, s, 0): present -> 1-based index, absent -> 0
        out = np.zeros(len(sel), dtype=np.int64)
        selc = np.ascontiguousarray(sel)
        orc.lib().orc_rank(m._c, orc._ptr(selc), len(selc), orc._ptr(out))
        assert np.array_equal(ref[out - 1], selc)
        if nup < L:
            absent = np.array([(1 << L) - 1], dtype=np.uint64)
            o = np.ones(1, dtype=np.int64)
            orc.lib().orc_rank(m._c, orc._ptr(absent), 1, orc._ptr(o))
            assert o[0] == 0


def test_model_structure():
    """test_SpinModel.jl:9-48."""
    m = orc.build_model(4)
    assert m.mode == "full" and len(m) == 16 and m.nup is None
    ms = orc.build_model(4, nup=2)
    assert ms.mode == "sector" and all(bin(int(s)).count("1") == 2 for s in ms.states)
    assert orc.nn_hopping(4, 1.0) == [(1, 2, 1.0), (2, 3, 1.0), (3, 4, 1.0)]
    lr = orc.long_range_hopping(3, lambda i, j: 1.0 / abs(i - j))
    assert lr == [(1, 2, 1.0), (1, 3, 0.5), (2, 3, 1.0)]
    x = orc.XXZChain(6, nup=3)
    assert np.allclose(orc.momenta(x), 2 * np.pi * np.arange(6) / 6)
    xp = orc.XXZChain(4, boundary="periodic")
    assert (4, 1, 0.5) in xp.hopping_list and (4, 1, 1.0) in xp.zz_list
    x2 = orc.XXZChain(2, boundary="periodic", nup=1)
    assert len(x2.hopping_list) == 1                    # ring closed only if L > 2
    with pytest.raises(ValueError):
        orc.XXZChain(4, boundary="twisted")


# --------------------------------------------------------------- apply_H!

def test_xxz_L2_known_matrix():
    """test_PublicAPI.jl:5-28."""
    m = orc.XXZChain(2, Jxy=1.0, Jz=1.0, nup=1)
    H = np.zeros((2, 2))
    for j in range(2):
        e = np.zeros(2)
        e[j] = 1.0
        out = np.zeros(2)
        orc.apply_H_(out, e, m)
        H[:, j] = out
    assert np.allclose(H, [[-0.25, 0.5], [0.5, -0.25]], atol=0, rtol=1e-15)
    assert np.allclose(np.linalg.eigvalsh(H), [-0.75, 0.25])


def _random_lists(L, rng, long_range):
    if long_range:
        hop = [(i, j, rng.normal()) for i in range(1, L + 1) for j in range(i + 1, L + 1)
               if rng.random() < 0.5]
        zz = [(j, i, rng.normal()) for i in range(1, L + 1) for j in range(i + 1, L + 1)
              if rng.random() < 0.5]
    else:
        hop = [(i, i + 1, rng.normal()) for i in range(1, L)]
        zz = [(i, i + 1, rng.normal()) for i in range(1, L)]
    return hop, zz, rng.normal(size=L)


@pytest.mark.parametrize("L,nup", [(2, 1), (4, None), (4, 2), (5, 2), (6, 3), (7, None),
                                   (8, 4), (9, 3), (10, 5), (10, 0), (10, 10)])
@pytest.mark.parametrize("kind", ["xxz_open", "xxz_periodic", "random_nn", "random_lr"])
def test_apply_H_vs_independent_dense(L, nup, kind):
    rng = np.random.default_rng(L * 131 + (nup or 77))
    if kind == "xxz_open":
        hop, zz, fld = dense_ref.xxz_lists(L, 0.8, 1.3, 0.2)
    elif kind == "xxz_periodic":
        hop, zz, fld = dense_ref.xxz_lists(L, 1.0, -0.7, 0.0, periodic=True)
    else:
        hop, zz, fld = _random_lists(L, rng, kind == "random_lr")
    m = orc.build_model(L, nup=nup, hopping=hop, onsite_field=fld, zz=zz)
    Hd = dense_ref.dense_H(L, nup, hop, zz, fld)
    N = len(m)
    assert Hd.shape == (N, N)
    psi = rng.normal(size=N)
    out_c = np.empty(N)
    out_np = np.empty(N)
    orc.apply_H_(out_c, psi, m)
    orc.apply_H_np_(out_np, psi, m)
    ref = Hd @ psi
    scale = max(np.linalg.norm(ref), 1e-300)
    assert np.linalg.norm(out_c - ref) / scale < 1e-14
    assert np.linalg.norm(out_np - ref) / scale < 1e-14
    psic = rng.normal(size=N) + 1j * rng.normal(size=N)
    outc = np.empty(N, dtype=np.complex128)
    orc.apply_H_(outc, psic, m)
    refc = Hd @ psic
    assert np.linalg.norm(outc - refc) / max(np.linalg.norm(refc), 1e-300) < 1e-14


def test_apply_rescaled_and_szq():
    """Hamiltonian.jl:286-301; test_Hamiltonian.jl:93-110."""
    m = orc.XXZChain(6, nup=3)
    N = len(m)
    psi = RNG.normal(size=N)
    Hp = np.empty(N)
    orc.apply_H_(Hp, psi, m)
    out = np.empty(N)
    orc.apply_rescaled_H_(out, psi, orc.apply_H_, m, 2.5, -0.3)
    assert np.allclose(out, (Hp - (-0.3) * psi) / 2.5, rtol=1e-15, atol=1e-16)
    out2 = np.empty(N)
    orc.lib().orc_apply_rescaled_H(m._c, orc._ptr(out2), orc._ptr(psi), 2.5, -0.3, 0)
    assert np.array_equal(out, out2)

    q = np.pi / 3
    phi = orc.Sz_q_vector(m, psi, q)
    st = m.states
    phi_ref = np.zeros(N, dtype=np.complex128)
    for r in range(1, m.L + 1):           # sum_r e^{iq(r-1)}/sqrt(L) Sz_r psi
        szr = np.where((st >> np.uint64(r - 1)) & np.uint64(1) == 1, 0.5, -0.5)
        phi_ref += np.exp(1j * q * (r - 1)) / np.sqrt(m.L) * szr * psi
    assert np.allclose(phi, phi_ref, atol=1e-12, rtol=0)
    assert np.allclose(orc.Sz_q_vector_np(m, psi, q), phi, atol=1e-14, rtol=0)
    psic = psi + 1j * RNG.normal(size=N)
    assert np.allclose(orc.Sz_q_vector(m, psic, q), orc.Sz_q_vector_np(m, psic, q), atol=1e-14)


# ---------------------------------------------------------------- Lanczos

HEIS_E0 = {4: -1.616025403784439, 6: -2.493577133887925, 8: -3.374932598687896,
           10: -4.258035207282884, 12: -5.1420906328405}


@pytest.mark.parametrize("L", [4, 6, 8, 10, 12])
def test_groundstate_known_energies(L):
    """open Heisenberg chain E0 (independent dense/ARPACK values)."""
    m = orc.XXZChain(L, nup=L // 2)
    E0, psi = orc.groundstate(m, lanc_m=min(len(m), 120), rng=np.random.default_rng(L))
    assert abs(E0 - HEIS_E0[L]) < 1e-10
    Hp = np.empty_like(psi)
    orc.apply_H_(Hp, psi, m)
    assert np.linalg.norm(Hp - E0 * psi) < 1e-8
    if L <= 10:
        hop, zz, fld = dense_ref.xxz_lists(L)
        ev = np.linalg.eigvalsh(dense_ref.dense_H(L, L // 2, hop, zz, fld))
        assert abs(ev[0] - HEIS_E0[L]) < 1e-11
        assert abs(ev[-1] - (L - 1) / 4) < 1e-12


def test_lanczos_vs_exact_diag_L6():
    """test_Lanczos.jl:29-54 (lanc_m = N, atol 1e-12; residual < 1e-10)."""
    m = orc.XXZChain(6, nup=3)
    N = len(m)
    hop, zz, fld = dense_ref.xxz_lists(6)
    H = dense_ref.dense_H(6, 3, hop, zz, fld)
    E0, psi = orc.groundstate(m, lanc_m=N, rng=np.random.default_rng(3))
    assert abs(E0 - np.linalg.eigvalsh(H)[0]) < 1e-12
    assert abs(np.linalg.norm(psi) - 1) < 1e-12
    assert np.linalg.norm(H @ psi - E0 * psi) < 1e-10


def test_lanczos_capping_and_extremal():
    """test_Lanczos.jl:57-119."""
    m = orc.XXZChain(4, nup=2)
    N = len(m)
    hop, zz, fld = dense_ref.xxz_lists(4)
    ev = np.linalg.eigvalsh(dense_ref.dense_H(4, 2, hop, zz, fld))
    E0, psi = orc.groundstate(m, lanc_m=100, rng=np.random.default_rng(5))
    assert len(psi) == N and abs(np.linalg.norm(psi) - 1) < 1e-12
    Emin, Emax = orc.lanczos_extremal(orc.apply_H_, m, lanc_m=100, rng=np.random.default_rng(6))
    assert abs(Emin - ev[0]) < 1e-12 and abs(Emax - ev[-1]) < 1e-12
    v = orc.randn_complex(np.random.default_rng(7), N)
    v /= np.linalg.norm(v)
    a, b, _ = orc.lanczos_tridiag(orc.apply_H_, m, v, lanc_m=100)
    assert len(a) <= N and len(b) == len(a) - 1
    with pytest.raises(RuntimeError):
        orc.lanczos_tridiag(orc.apply_H_, m, np.zeros(N, dtype=complex))


def test_lanczos_tridiag_complex_alpha():
    """test_Lanczos.jl:6-26."""
    m = orc.XXZChain(2, nup=1)
    v = np.array([1.0, 1j]) / np.sqrt(2)
    Hv = np.empty_like(v)
    orc.apply_H_(Hv, v, m)
    a, b, normv = orc.lanczos_tridiag(orc.apply_H_, m, v, lanc_m=2)
    assert abs(a[0] - np.vdot(v, Hv).real) < 1e-12 and abs(normv - 1) < 1e-12


def test_lanczos_reproducible():
    """test_Lanczos.jl:122-166."""
    m = orc.XXZChain(6, nup=3)
    E1, p1 = orc.groundstate(m, lanc_m=10, rng=np.random.default_rng(1234))
    E2, p2 = orc.groundstate(m, lanc_m=10, rng=np.random.default_rng(1234))
    assert abs(E1 - E2) < 1e-14 and np.allclose(p1, p2, atol=1e-14)
    b1 = orc.lanczos_extremal(orc.apply_H_, m, lanc_m=10, rng=np.random.default_rng(42))
    b2 = orc.lanczos_extremal(orc.apply_H_, m, lanc_m=10, rng=np.random.default_rng(42))
    assert np.allclose(b1, b2, atol=1e-14)


# ---------------------------------------------------------- time evolution

def test_time_evolve_L2_known_answers():
    """test_PublicAPI.jl:56-134; SURVEY 8c psi(0.3) value."""
    m = orc.XXZChain(2, nup=1)
    H = np.array([[-0.25, 0.5], [0.5, -0.25]])
    psi0 = np.array([1.0 + 0j, 0.0])
    exact = expm(-0.3j * H) @ psi0
    assert np.allclose(exact, [0.98599146 + 0.07408833j, 0.01119736 - 0.14901803j], atol=1e-8)
    pk = orc.time_evolve(m, psi0, 0.3, method="krylov", kry_m=2)
    assert np.allclose(pk, exact, atol=1e-10) and abs(np.linalg.norm(pk) - 1) < 1e-12
    assert np.allclose(orc.time_evolve(m, psi0, 0.0, method="krylov", kry_m=2), psi0, atol=1e-12)
    pc = orc.time_evolve(m, psi0, 0.3, method="chebyshev", cheb_n=30, Ebounds=(-0.75, 0.25))
    assert np.allclose(pc, exact, atol=1e-8) and abs(np.linalg.norm(pc) - 1) < 1e-8
    pa = orc.time_evolve(m, psi0, 0.1, method="chebyshev", cheb_n=20)
    assert abs(np.linalg.norm(pa) - 1) < 1e-6
    with pytest.raises(ValueError):
        orc.time_evolve(m, psi0, 0.3, method="unknown")


@pytest.mark.parametrize("method", ["krylov", "chebyshev"])
def test_time_evolve_vs_expm_L8(method):
    L, nup = 8, 4
    m = orc.XXZChain(L, nup=nup, Jz=0.6)
    hop, zz, fld = dense_ref.xxz_lists(L, 1.0, 0.6)
    H = dense_ref.dense_H(L, nup, hop, zz, fld)
    psi0 = orc.neel_state(m).astype(np.complex128)
    exact = expm(-0.5j * H) @ psi0
    if method == "krylov":
        p = orc.time_evolve(m, psi0, 0.5, method="krylov", kry_m=30)
    else:
        ev = np.linalg.eigvalsh(H)
        p = orc.time_evolve(m, psi0, 0.5, method="chebyshev", cheb_n=40, Ebounds=(ev[0], ev[-1]))
    assert np.allclose(p, exact, atol=1e-9)


# ---------------------------------------------------------------- S(q, w)

def test_kpm_sum_rule():
    """test_KPM.jl:67-91 (rtol 5e-3); SURVEY 8c exact weight 0.7052153237226118."""
    m = orc.XXZChain(6, nup=3)
    _, psi0 = orc.groundstate(m, lanc_m=20, rng=np.random.default_rng(11))
    q = np.pi
    w = np.arange(0.0, 5.0 + 1e-12, 0.01)
    phi = orc.Sz_q_vector(m, psi0, q)
    exact_weight = np.linalg.norm(phi) ** 2
    assert abs(exact_weight - 0.7052153237226118) < 1e-9
    S = orc.dynamical_structure_factor(m, psi0, [q], w, method="kpm", kpm_m=120,
                                       kernel="jackson", rng=np.random.default_rng(12))
    assert np.all(np.isfinite(S)) and np.all(S >= 0)
    assert abs(S[0].sum() * (w[1] - w[0]) - exact_weight) < 5e-3 * exact_weight


def test_kpm_rescaling():
    """test_KPM.jl:4-41."""
    a, b = orc._rescaling_from_bounds(-3.0, 5.0)
    assert abs((-3 - b) / a + 0.99) < 1e-12 and abs((5 - b) / a - 0.99) < 1e-12
    m = orc.XXZChain(6, nup=3)
    Emin, Emax = orc.estimate_energy_bounds(orc.apply_H_, m, lanc_m=20, rng=np.random.default_rng(1))
    a, b = orc._rescaling_from_bounds(Emin, Emax)
    assert (Emin - b) / a > -1 and (Emax - b) / a < 1


def test_lanczos_sqw_shape_and_sign():
    """test_PublicAPI.jl:154-183."""
    m = orc.XXZChain(4, nup=2)
    _, psi0 = orc.groundstate(m, lanc_m=6, rng=np.random.default_rng(2))
    q = orc.momenta(m)
    w = np.linspace(0, 3, 40)
    S = orc.dynamical_structure_factor(m, psi0, q, w, method="lanczos", lanc_m=6, eta=0.05)
    assert S.shape == (4, 40) and np.all(np.isfinite(S)) and np.all(S >= -1e-12)
    with pytest.raises(ValueError):
        orc.dynamical_structure_factor(m, psi0, q, w, method="unknown")
    # spectral weight: sum over poles equals ||phi||^2 (Lorentzians integrate to 1)
    wide = np.linspace(-40, 40, 160001)
    Sw = orc.lanczos_sqw(psi0, m, [np.pi], wide, lanc_m=6, eta=0.05)
    phi = orc.Sz_q_vector(m, psi0, np.pi)
    assert abs(Sw[0].sum() * (wide[1] - wide[0]) - np.linalg.norm(phi) ** 2) < 2e-3


# ---------------------------------------------------------- initial states

def test_initial_states():
    """test_InitialStates.jl:6-108."""
    mf = orc.build_model(4)
    assert np.argmax(orc.neel_state(mf)) == 0b0101
    assert orc.polarized_state(mf, up=True)[15] == 1.0
    assert orc.polarized_state(mf, up=False)[0] == 1.0
    assert np.argmax(orc.polarized_state_with_flips(mf, [2, 4])) == 0b0101
    assert np.argmax(orc.domain_wall_state(mf)) == 0b0011
    ms = orc.build_model(4, nup=2)
    for f in (orc.neel_state, orc.domain_wall_state):
        v = f(ms)
        assert v.dtype == np.float64 and v.sum() == 1.0 and np.count_nonzero(v) == 1
    assert ms.states[np.argmax(orc.neel_state(ms))] == 0b0101
    with pytest.raises(ValueError):
        orc.polarized_state(ms, up=True)
    with pytest.raises(ValueError):
        orc.polarized_state_with_flips(ms, [1])
    with pytest.raises(ValueError):
        orc.polarized_state_with_flips(mf, [5])


def test_seeded_row_matches_full_apply():
    """The sampled-row checker used at L>=32 agrees with the full oracle apply."""
    L, nup, seed = 12, 6, 20261018
    m = orc.XXZChain(L, nup=nup)
    N = len(m)
    psi = orc.fill_seeded(N, seed)
    out = np.empty(N)
    orc.apply_H_(out, psi, m)
    args = (L, nup, m.hopping_list, m.zz_list, m.onsite_field)
    for idx in RNG.integers(0, N, 50):
        s = int(m.states[idx])
        assert abs(orc.row_seeded_f64(args, s, seed) - out[idx]) < 1e-14


def test_energy_golden_file_agrees_with_the_oracle_lanczos():
    """tests/golden/energy_golden.json (ARPACK on the oracle matvec, make_energy_golden.py) vs the
    oracle's own Lanczos.jl restatement at the one size that runs in seconds; pins the fixture that
    the full-size GPU ground-state tests compare with."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "energy_golden.json")))
    assert set(g["E0"]) == {"18", "20", "22", "24"} and max(g["residual"].values()) < 1e-12
    L = 18
    m = orc.XXZChain(L, nup=L // 2)
    E0, psi = orc.groundstate(m, lanc_m=90, rng=np.random.default_rng(L))
    assert abs(E0 - g["E0"][str(L)]) < 1e-10
    e = np.array([g["E0"][str(x)] for x in (18, 20, 22, 24)])
    assert np.all(np.abs(np.diff(e, 2)) < 2e-4)          # E0(L) is almost linear in L (bulk energy density)


@pytest.mark.parametrize("L,nup,boundary", [(12, 6, "open"), (13, 4, "open"), (10, 5, "periodic"), (9, None, "open"), (8, 0, "open")])
def test_ranked_cpu_baseline_equals_the_reference_loop(L, nup, boundary):
    """bench.py's second CPU figure (combinatorial ranking instead of the Dict probe) computes the
    same H.psi as the reference-faithful loop, also for long-range hops."""
    rng = np.random.default_rng(L)
    m = orc.XXZChain(L, Jxy=0.8, Jz=1.1, hz=0.3, nup=nup, boundary=boundary)
    psi = rng.standard_normal(len(m))
    a, b = np.empty_like(psi), np.empty_like(psi)
    orc.apply_H_(a, psi, m)
    orc.apply_H_ranked_(b, psi, m)
    assert np.array_equal(a, b)
    if nup is not None and 0 < nup < L:
        hop = [(1, L, 0.4), (2, 5, -0.7), (L - 1, 3, 0.2)] + [(i, i + 1, 0.5) for i in range(1, L)]
        m2 = orc.build_model(L, nup=nup, hopping=hop, onsite_field=np.zeros(L), zz=[])
        orc.apply_H_(a, psi, m2)
        orc.apply_H_ranked_(b, psi, m2)
        assert np.array_equal(a, b)
