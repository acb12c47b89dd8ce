"""Host-side logic of the Python mirror that needs no GPU: argument checks that mirror the reference's assertions
(raised before any library call) and the in-place wrappers, with the device routine replaced by a stub."""
import numpy as np
import pytest

from conftest import sd
from spindyn import api


class _M:
    dim = 4


def test_inplace_krylov_wrapper_logic(monkeypatch):
    """krylov_time_evolve! (Krylov.jl:55-118): size assertions, None on success, psi_out on the zero-norm early
    return (:69-72), ComplexF64-only method."""
    def fake(psi0, dt, applyH_, model, kry_m=30, device=False):
        if not np.any(psi0):
            return np.array(psi0, copy=True)
        return psi0 * np.exp(-1j * dt)

    monkeypatch.setattr(api, "krylov_time_evolve", fake)
    out, psi = np.zeros(4, dtype=complex), np.arange(4) + 0j
    ws = sd.KrylovWorkspace(4, 20)
    assert api.krylov_time_evolve_(out, psi, 0.3, sd.apply_H_, _M(), ws, kry_m=20) is None
    assert np.allclose(out, psi * np.exp(-0.3j))
    assert api.krylov_time_evolve_(out, np.zeros(4, dtype=complex), 0.3, sd.apply_H_, _M(), ws, kry_m=20) is out
    assert not np.any(out)
    with pytest.raises(ValueError):
        api.krylov_time_evolve_(np.zeros(3, dtype=complex), psi, 0.3, sd.apply_H_, _M(), ws, kry_m=20)
    with pytest.raises(ValueError):
        api.krylov_time_evolve_(out, psi, 0.3, sd.apply_H_, _M(), sd.KrylovWorkspace(4, 5), kry_m=20)
    with pytest.raises(TypeError):
        api.krylov_time_evolve_(np.zeros(4), psi, 0.3, sd.apply_H_, _M(), ws)


def test_chebyshev_argument_checks_precede_the_library():
    """Chebyshev.jl:66 (`cheb_n >= 1`), :86 (workspace size), and the callback restriction of the GPU path."""
    psi = np.arange(4) + 0j
    with pytest.raises(ValueError):
        sd.chebyshev_time_evolve(psi, 0.1, sd.apply_H_, _M(), cheb_n=0)
    with pytest.raises(ValueError):
        sd.chebyshev_time_evolve(psi, 0.1, sd.apply_H_, _M(), cheb_n=5, workspace=sd.ChebyshevWorkspace(3))
    with pytest.raises(NotImplementedError):
        sd.chebyshev_time_evolve(psi, 0.1, lambda o, p, m: o, _M(), cheb_n=5)
    assert sd.ChebyshevWorkspace(psi).N == 4 and sd.ChebyshevWorkspace(7).N == 7


def test_chebyshev_coefficients_match_the_reference_formula():
    """Chebyshev.jl:70-79: a = (Emax-Emin)/(2*0.9999), b = (Emax+Emin)/2, c_k = (2-delta_k0)(-i)^k J_k(a dt) e^{-i b dt}."""
    from scipy.special import jv
    c, a, b = sd.chebyshev_coefficients(0.3, 6, (-0.75, 0.25))
    assert abs(a - 1.0 / (2 * 0.9999)) < 1e-15 and abs(b + 0.25) < 1e-15
    want = [(1 if k == 0 else 2) * (-1j) ** k * jv(k, a * 0.3) * np.exp(-1j * b * 0.3) for k in range(6)]
    assert np.allclose(c, want, rtol=0, atol=1e-15)


def test_threaded_q_loop_partitions_q_in_order(monkeypatch):
    """_q_parallel (the reference's Threads.@threads q-loop, LanczosSqw.jl:65 / KPM_Sqw.jl:218): contiguous shares of
    q_list, one context + model copy per thread, rows concatenated in q order, context closed after its model."""
    events = []

    class Ctx:
        device, world = 0, 1

        def __init__(self, d=0):
            events.append("ctx")

        def close(self):
            events.append("close")

    class Mdl:
        def __init__(self, L, nup, h, f, z, ctx=None):
            self.L, self.nup, self.hopping_list, self.onsite_field, self.zz_list, self.ctx = L, nup, h, f, z, ctx or Ctx()

        def __del__(self):
            events.append("model_del")

    monkeypatch.setattr(api, "Context", Ctx)
    monkeypatch.setattr(api, "Model", Mdl)

    def fn(psi, m, qs, q_threads=1, **kw):
        assert q_threads == 1 and isinstance(psi, np.ndarray)
        return np.array([[q * 10 + kw["x"]] for q in qs])

    m = Mdl(4, 2, [], [0.0] * 4, [])
    events.clear()
    out = api._q_parallel(fn, np.zeros(3), m, [0, 1, 2, 3, 4, 5, 6], 3, x=1)
    assert out.ravel().tolist() == [1, 11, 21, 31, 41, 51, 61]
    assert events.count("ctx") == 3 and events.count("close") == 3 and events.count("model_del") == 3
    for i, e in enumerate(events):
        if e == "close":
            assert "model_del" in events[:i]
    assert api._q_parallel(fn, np.zeros(3), m, [5], 8, x=2).ravel().tolist() == [52]


@pytest.mark.parametrize("kernel", ["jackson", "lorentz", "none"])
def test_kpm_reconstruction_equals_the_oracles(kernel):
    """The host half of kpm_sw / kpm_sqw (KPM_Sqw.jl:48-92: kernel damping, x = (w + E0 - b) / a, the Chebyshev series,
    1 / (a pi sqrt(1 - x^2)), max(0, .)) shared by the per-momentum and the q-batched paths, against the oracle's loop on
    the same moments -- including frequencies outside the rescaled band (S = 0)."""
    import oracle.oracle as orc
    om = orc.XXZChain(8, nup=4)
    rng = np.random.default_rng(11)
    phi = rng.standard_normal(len(om)) + 1j * rng.standard_normal(len(om))
    phi /= np.linalg.norm(phi)
    a, b, E0, M = 2.9, -0.3, -3.1, 37
    w = np.linspace(-1.0, 7.0, 41)
    mu = orc.compute_chebyshev_moments(orc.apply_H_, phi, M, a, b, om)
    want = orc.kpm_sw(phi, orc.apply_H_, om, w, a, b, E0, kpm_m=M, kernel=kernel)
    got = api._kpm_reconstruct(np.asarray(mu), w, a, b, E0, M, kernel)
    assert np.any(want == 0.0) and np.any(want > 0.0)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-14)


def test_q_batch_automatic_choice():
    """q_batch=None batches small bases on a single GPU only (the multi-vector kernel loses to the block kernel at large N)."""
    class Ctx:
        world = 1

    class Mdl:
        ctx = Ctx()
        dim = 12870

    m = Mdl()
    assert api._use_q_batch(m, [0.0, 1.0], 1, None) is True
    assert api._use_q_batch(m, [0.0], 1, None) is False                 # one momentum: nothing to batch
    assert api._use_q_batch(m, [0.0, 1.0], 4, None) is False            # q_threads asks for the threaded loop
    m.dim = api.Q_BATCH_AUTO_MAX_DIM + 1
    assert api._use_q_batch(m, [0.0, 1.0], 1, None) is False
    assert api._use_q_batch(m, [0.0, 1.0], 1, True) is True             # explicit request wins
    m.ctx.world = 2
    assert api._use_q_batch(m, [0.0, 1.0], 1, None) is False
    with pytest.raises(ValueError):
        api._use_q_batch(m, [0.0, 1.0], 1, True)
