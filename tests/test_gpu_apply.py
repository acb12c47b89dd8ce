"""GPU parity tests of the H.psi path: every call goes through the C ABI of
libspindyn_cuda (via the ctypes mirror) and is compared with the CPU oracle on
the same seeded inputs.  Tolerance: 1e-13 relative L2 for H.psi (north star);
basis enumeration and ranks bit-exact."""
import ctypes
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.oracle as orc  # noqa: E402
from conftest import sd  # noqa: E402

TOL = 1e-13


def rel(a, b):
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (n if n > 0 else 1.0)


def rand_vec(rng, n, dtype):
    v = rng.standard_normal(n)
    if np.dtype(dtype) == np.complex128:
        v = v + 1j * rng.standard_normal(n)
    return v.astype(dtype)


def paths_of(m):
    """Every kernel path the model qualifies for ("block" stores vectors in block layout, so a
    path switch is only legal while no DeviceVector of the model is alive)."""
    out = []
    for p in ("block", "tiled", "generic"):
        try:
            m.set_path(p)
            out.append(p)
        except NotImplementedError:
            pass
    return out


def both(L, nup=None, Jxy=1.0, Jz=1.0, hz=0.0, boundary="open"):
    return (sd.XXZChain(L, Jxy=Jxy, Jz=Jz, hz=hz, nup=nup, boundary=boundary),
            orc.XXZChain(L, Jxy=Jxy, Jz=Jz, hz=hz, nup=nup, boundary=boundary))


# ---------------------------------------------------------------- basis

@pytest.mark.parametrize("L,nup", [(1, 0), (1, 1), (4, 2), (6, 3), (10, 5), (12, 4), (16, 8), (20, 10), (17, 3)])
def test_sector_basis_bit_exact(L, nup):
    """build_sector_basis order (Basis.jl:37-53) == device unranking, bit for bit."""
    states, idxmap = sd.build_sector_basis(L, nup)
    ref = np.zeros(orc.lib().orc_sector_dim(L, nup), dtype=np.uint64)
    orc.lib().orc_build_sector_basis(L, nup, orc._ptr(ref))
    assert states.dtype == np.uint64 and np.array_equal(states, ref)
    m = sd.build_model(L, nup=nup)
    idx = m.rank_of(ref)
    assert np.array_equal(idx, np.arange(1, len(ref) + 1))          # idxmap[s] = i (1-based)


def test_basis_edge_cases_reference_tests():
    """test_Basis.jl:4-19."""
    for bad in [(0, None), (-1, None)]:
        with pytest.raises(ValueError):
            sd.build_full_basis(bad[0])
    with pytest.raises(ValueError):
        sd.build_sector_basis(4, -1)
    with pytest.raises(ValueError):
        sd.build_sector_basis(4, 5)
    with pytest.raises(ValueError):
        sd.build_full_basis(64)
    st, mp = sd.build_sector_basis(4, 0)
    assert list(st) == [0] and mp[0] == 1
    st, mp = sd.build_sector_basis(4, 4)
    assert len(st) == 1 and bin(int(st[0])).count("1") == 4 and mp[int(st[0])] == 1
    st, mp = sd.build_full_basis(4)
    assert np.array_equal(st, np.arange(16, dtype=np.uint64)) and mp[5] == 6


def test_rank_absent_states():
    m = sd.build_model(8, nup=3)
    q = np.array([0b111, 0b1111, 1 << 8 | 0b11, 0b10101000, 0], dtype=np.uint64)
    r = m.rank_of(q)
    assert r[0] == 1 and r[1] == 0 and r[2] == 0 and r[3] > 0 and r[4] == 0


def test_unrank_window_large_u64_ranks():
    """L=36 nup=18 (9.08e9 states, ranks beyond 2^32): windows of the basis against
    the oracle's closed-form rank; round trip through sd_rank."""
    L, nup = 36, 18
    m = sd.build_model(L, nup=nup)
    assert m.dim == 9075135300 and m.info["rank_bits"] == 64
    for first in [0, 2 ** 32 - 5, 5_000_000_000, m.dim - 1000]:
        st = m.unrank(first, 1000)
        ranks = np.array([orc.lib().orc_rank_closed_form(L, nup, int(s)) for s in st], dtype=np.uint64)
        assert np.array_equal(ranks, np.arange(first, first + 1000, dtype=np.uint64))
        assert np.array_equal(m.rank_of(st), np.arange(first + 1, first + 1001))


# ---------------------------------------------------------------- apply_H

CASES = [
    # L, nup, boundary
    (2, 1, "open"), (4, 2, "open"), (6, 3, "open"), (6, None, "open"), (8, None, "periodic"),
    (10, 5, "open"), (12, 6, "open"), (12, 3, "open"), (13, 6, "open"), (14, 7, "open"), (14, 7, "periodic"),
    (16, 8, "open"), (16, 2, "open"), (16, 14, "open"), (18, 9, "open"), (20, 10, "open"), (15, 0, "open"),
    (15, 15, "open"), (12, None, "open"),
    (16, 8, "periodic"), (17, 5, "periodic"), (18, 9, "periodic"), (16, 15, "periodic"),   # block kernel + wrap pass
]


@pytest.mark.parametrize("L,nup,boundary", CASES)
@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_apply_H_matches_oracle(L, nup, boundary, dtype):
    m, om = both(L, nup, Jxy=0.7, Jz=1.3, hz=0.2, boundary=boundary)
    rng = np.random.default_rng(L * 100 + (nup or 0))
    psi = rand_vec(rng, m.dim, dtype)
    ref = np.empty_like(psi)
    orc.apply_H_(ref, psi, om)
    for path in paths_of(m):
        m.set_path(path)
        out = np.full_like(psi, np.nan)
        assert sd.apply_H_(out, psi, m) is out
        assert rel(out, ref) < TOL, (path, rel(out, ref))


def test_tiled_path_selected_for_open_chain():
    assert sd.XXZChain(16, nup=8).info["kernel_path"] == "block"
    assert sd.XXZChain(14, nup=7).info["kernel_path"] == "tiled"
    assert paths_of(sd.XXZChain(20, nup=10)) == ["block", "tiled", "generic"]
    with pytest.raises(NotImplementedError):
        sd.XXZChain(14, nup=7).set_path("block")
    assert sd.XXZChain(16, nup=8, boundary="periodic").info["kernel_path"] == "block"    # wrap pass + block kernel
    assert paths_of(sd.XXZChain(16, nup=8, boundary="periodic")) == ["block", "generic"]    # the tiled kernel has no wrap bond
    assert sd.XXZChain(14, nup=7, boundary="periodic").info["kernel_path"] == "generic"
    assert sd.XXZChain(12).info["kernel_path"] == "generic"
    with pytest.raises(NotImplementedError):
        sd.XXZChain(12).set_path("tiled")


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_apply_H_random_couplings_and_long_range(dtype):
    """Arbitrary per-bond J / Jz / per-site field (tiled) and long-range lists (generic)."""
    rng = np.random.default_rng(7)
    for (L, nup, default) in [(14, 6, "tiled"), (17, 8, "block"), (18, 5, "block")]:
        hop = [(i, i + 1, rng.uniform(0.2, 1.5)) for i in range(1, L)]
        zz = [(i, i + 1, rng.uniform(-1, 1)) for i in range(1, L)]
        fld = rng.uniform(-1, 1, L)
        m = sd.build_model(L, nup=nup, hopping=hop, onsite_field=fld, zz=zz)
        om = orc.build_model(L, nup=nup, hopping=hop, onsite_field=fld, zz=zz)
        assert m.info["kernel_path"] == default
        psi = rand_vec(rng, m.dim, dtype)
        ref = np.empty_like(psi)
        orc.apply_H_(ref, psi, om)
        for path in paths_of(m):
            m.set_path(path)
            out = np.empty_like(psi)
            sd.apply_H_(out, psi, m)
            assert rel(out, ref) < TOL, (L, nup, path)
    L, nup = 14, 6
    fld = rng.uniform(-1, 1, L)
    hop = sd.long_range_hopping(L, lambda i, j: 1.0 / abs(i - j) ** 2)
    zz = [(i, j, 0.5 / abs(i - j)) for i in range(1, L + 1) for j in range(i + 1, L + 1)]
    for nup_ in (nup, None):
        m = sd.build_model(L, nup=nup_, hopping=hop, onsite_field=fld, zz=zz)
        om = orc.build_model(L, nup=nup_, hopping=hop, onsite_field=fld, zz=zz)
        assert m.info["kernel_path"] == "generic"
        psi = rand_vec(rng, m.dim, dtype)
        ref = np.empty_like(psi)
        orc.apply_H_(ref, psi, om)
        out = np.empty_like(psi)
        sd.apply_H_(out, psi, m)
        assert rel(out, ref) < TOL


def test_reference_L2_matrix():
    """test_PublicAPI.jl:5-28: XXZChain(2, nup=1) is [[-1/4, 1/2], [1/2, -1/4]]."""
    m = sd.XXZChain(2, Jxy=1.0, Jz=1.0, hz=0.0, nup=1)
    H = np.zeros((2, 2))
    for j in range(2):
        e = np.zeros(2)
        e[j] = 1.0
        col = np.zeros(2)
        sd.apply_H_(col, e, m)
        H[:, j] = col
    assert np.allclose(H, [[-0.25, 0.5], [0.5, -0.25]], atol=1e-15)


def test_apply_H_argument_errors():
    m = sd.XXZChain(6, nup=3)
    psi = np.zeros(m.dim)
    with pytest.raises(ValueError):
        sd.apply_H_(np.zeros(m.dim + 1), psi, m)
    with pytest.raises(TypeError):
        sd.apply_H_(np.zeros(m.dim, dtype=np.complex128), psi, m)
    d = m.to_device(psi)
    with pytest.raises(ValueError):
        sd.apply_H_(d, d, m)                         # out must not alias psi


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_rescaled_and_cheb_step_fusions(dtype):
    """apply_rescaled_H! (Hamiltonian.jl:286-301) and the fused Chebyshev step
    (KPM_Sqw.jl:111-117, Chebyshev.jl:112-116) against numpy on oracle H.psi."""
    a, b = 3.7, -0.4
    rng = np.random.default_rng(3)
    for (L, nup, path) in [(14, 7, "tiled"), (14, 7, "generic"), (17, 8, "block"), (16, 9, "block")]:
        m, om = both(L, nup, Jz=0.8)
        m.set_path(path)
        v = rand_vec(rng, m.dim, dtype)
        vprev = rand_vec(rng, m.dim, dtype)
        phi = rand_vec(rng, m.dim, dtype)
        acc0 = rand_vec(rng, m.dim, dtype)
        Hv = np.empty_like(v)
        orc.apply_H_(Hv, v, om)
        ref_resc = (Hv - b * v) / a
        out = np.empty_like(v)
        sd.apply_rescaled_H_(out, v, sd.apply_H_, m, a, b)
        assert rel(out, ref_resc) < TOL
        ref_next = 2.0 * ref_resc - vprev
        ck = 0.3 - 0.2j if dtype == np.complex128 else 0.3
        dv, dprev, dphi, dacc = (m.to_device(x) for x in (v, vprev, phi, acc0))
        mu, n2 = ctypes.c_double(), ctypes.c_double()
        ckc = complex(ck)
        sd._lib.check(sd.lib().sd_cheb_step(m._h, dprev._h, dv._h, dprev._h, a, b, dphi._h,
                                            ctypes.byref(mu), ctypes.byref(n2), dacc._h,
                                            sd._lib.SdComplex(ckc.real, ckc.imag)))
        assert rel(dprev.to_host(), ref_next) < TOL                  # vnext aliases vprev
        assert abs(mu.value - np.vdot(phi, ref_next).real) < 1e-11 * max(1.0, abs(mu.value))
        assert abs(n2.value - np.vdot(ref_next, ref_next).real) < 1e-12 * n2.value
        assert rel(dacc.to_host(), acc0 + ck * ref_next) < TOL


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_apply_H_dot_fusion_and_determinism(dtype):
    L, nup = 16, 8
    m, om = both(L, nup)
    rng = np.random.default_rng(11)
    psi = rand_vec(rng, m.dim, dtype)
    ref = np.empty_like(psi)
    orc.apply_H_(ref, psi, om)
    d, o = m.to_device(psi), m.vector(dtype)
    vals = []
    for _ in range(3):
        r = sd._lib.SdComplex()
        sd._lib.check(sd.lib().sd_apply_H_dot(m._h, o._h, d._h, ctypes.byref(r)))
        vals.append((r.re, r.im))
    assert vals[0] == vals[1] == vals[2]                             # test_Lanczos.jl:122-166
    assert abs(complex(*vals[0]) - np.vdot(psi, ref)) < 1e-12 * abs(np.vdot(psi, ref))
    assert rel(o.to_host(), ref) < TOL


@pytest.mark.parametrize("psi_dtype", [np.float64, np.complex128])
def test_Sz_q_vector(psi_dtype):
    """test_Hamiltonian.jl:93-110 (L=6, q=pi/3) plus a larger sector against the oracle."""
    for L, nup in [(6, 3), (6, None), (14, 7)]:
        m, om = both(L, nup)
        rng = np.random.default_rng(5)
        psi0 = rand_vec(rng, m.dim, psi_dtype)
        q = np.pi / 3
        phi = sd.Sz_q_vector(m, psi0, q)
        assert phi.dtype == np.complex128
        assert np.allclose(phi, orc.Sz_q_vector_np(om, psi0, q), atol=1e-12)
        assert np.allclose(phi, orc.Sz_q_vector(om, psi0, q), atol=1e-13)


def test_blas1_and_conversions():
    m = sd.XXZChain(12, nup=6)
    rng = np.random.default_rng(2)
    x = rand_vec(rng, m.dim, np.complex128)
    y = rand_vec(rng, m.dim, np.complex128)
    dx, dy = m.to_device(x), m.to_device(y)
    assert abs(dx.dot(dy) - np.vdot(x, y)) < 1e-12
    assert abs(dx.dotu(dy) - np.sum(x * y)) < 1e-12
    assert abs(dx.norm() - np.linalg.norm(x)) < 1e-12
    dy.axpy(0.5 - 2j, dx)
    assert rel(dy.to_host(), y + (0.5 - 2j) * x) < 1e-15
    dx.scale(1j)
    assert rel(dx.to_host(), 1j * x) < 1e-15
    r = rand_vec(rng, m.dim, np.float64)
    dr = m.to_device(r)
    assert rel(dr.astype(np.complex128).to_host(), r.astype(np.complex128)) == 0
    with pytest.raises(ValueError):
        dr.scale(1j)                                                 # InexactError
    one = m.vector(np.float64).set_onehot(5).to_host()
    assert one[5] == 1.0 and one.sum() == 1.0
    f = m.vector(np.complex128).fill_seeded(42, 0.5).to_host()
    assert np.array_equal(f, 0.5 * orc.fill_seeded(m.dim, 42, cplx=True))


# ---------------------------------------------------------------- golden + large-size properties

def test_golden_fixture():
    """tests/golden/apply_golden.npz was written by tests/golden/make_golden.py from the oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "apply_golden.npz"))
    for key in [k for k in g.files if k.startswith("out_")]:
        _, L, nup, kind = key.split("_")
        L, nup = int(L), int(nup)
        m = sd.XXZChain(L, Jxy=float(g["Jxy"]), Jz=float(g["Jz"]), hz=float(g["hz"]), nup=nup)
        cplx = kind == "c128"
        psi = orc.fill_seeded(m.dim, int(g["seed"]), cplx=cplx)
        for path in paths_of(m):
            m.set_path(path)
            out = np.empty_like(psi)
            sd.apply_H_(out, psi, m)
            assert rel(out, g[key]) < TOL, (key, path)


@pytest.mark.parametrize("L,nup", [(24, 12), (28, 14)])
def test_large_sampled_rows_and_linearity(L, nup):
    """At sizes the oracle cannot hold in full: sampled rows of H.psi for the
    counter-based seeded psi (oracle regenerates any element), Hermiticity
    <x,Hy> = <Hx,y>, and agreement of the two kernel paths."""
    m = sd.XXZChain(L, nup=nup)
    seed = 20261018
    x = m.vector(np.float64).fill_seeded(seed)
    y = m.vector(np.float64).fill_seeded(seed + 7)
    hx, hy = m.vector(np.float64), m.vector(np.float64)
    sd.apply_H_(hx, x, m)
    sd.apply_H_(hy, y, m)
    lhs, rhs = x.dot(hy).real, hx.dot(y).real
    assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs))
    out = hx.to_host()
    rng = np.random.default_rng(L)
    rows = np.unique(np.concatenate([rng.integers(0, m.dim, 2000), [0, 1, m.dim - 1, m.dim // 2]]))
    st = np.array([m.unrank(int(r), 1)[0] for r in rows[:50]], dtype=np.uint64)
    hop, zz = om_lists(L)
    for r, s in zip(rows[:50], st):
        ref = orc.row_seeded_f64((L, nup, hop, zz, np.zeros(L)), int(s), seed)
        assert abs(out[r] - ref) <= 1e-13 * max(1.0, abs(ref)), (r, out[r], ref)
    assert m.info["kernel_path"] == "block"
    del hy, y
    for path in ("tiled", "generic"):                                # rank-ordered vectors: a second model
        m2 = sd.XXZChain(L, nup=nup)
        m2.set_path(path)
        x2 = m2.vector(np.float64).fill_seeded(seed)
        hg = m2.vector(np.float64)
        sd.apply_H_(hg, x2, m2)
        assert rel(hg.to_host(), out) < TOL, path
        del x2, hg, m2


def om_lists(L, Jxy=1.0, Jz=1.0):
    return [(i, i + 1, Jxy / 2) for i in range(1, L)], [(i, i + 1, Jz) for i in range(1, L)]
