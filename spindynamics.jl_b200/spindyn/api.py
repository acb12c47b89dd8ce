"""Python mirror of the reference's operator / solver / public interface for the
H.psi hot path, on top of libspindyn_cuda.  Same names (Julia `f!` -> `f_`),
same argument meaning, same error behaviour (ArgumentError -> ValueError), so
tests/ can read like the reference's own tests.  Citations are file:line under
/root/reference/src.

Vector arguments may be numpy arrays (host, like Julia Vectors: uploaded,
processed on the GPU, result downloaded) or `DeviceVector`s (stay in HBM).
All vector arithmetic runs in CUDA kernels; the host keeps exactly what the
north star keeps there: random start vectors, m x m tridiagonal eigenproblems,
Bessel coefficients, kernel damping and spectrum reconstruction.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Optional, Sequence

import numpy as np

from ._lib import SdComplex, check, lib
from .core import Context, DeviceVector, Model, VecSet, _ptr, default_context

# ------------------------------------------------------------------ Basis.jl


def _validate_basis_args(L, nup=None):
    """Basis.jl:9-20."""
    if not L >= 1:
        raise ValueError("L must be at least 1")
    if not L <= 63:
        raise ValueError("L must be at most 63 when using UInt64 basis states")
    if nup is not None and not (0 <= nup <= L):
        raise ValueError("nup must satisfy 0 <= nup <= L")


class _RankMap:
    """Stands in for Dict{UInt64,Int} (Basis.jl:49-52): lookups are computed by
    ranking on the device instead of hashing."""

    def __init__(self, model: Model):
        self._m = model

    def __getitem__(self, s):
        r = int(self._m.rank_of([s])[0])
        if r == 0:
            raise KeyError(s)
        return r

    def get(self, s, default=0):
        r = int(self._m.rank_of([s])[0])
        return r if r != 0 else default

    def __contains__(self, s):
        return int(self._m.rank_of([s])[0]) != 0

    def __len__(self):
        return self._m.dim


def build_full_basis(L: int, ctx: Optional[Context] = None):
    """Basis.jl:23-34 -> (states, idxmap)."""
    _validate_basis_args(L)
    m = Model(L, None, [], np.zeros(L), [], ctx)
    return m.states, _RankMap(m)


def build_sector_basis(L: int, nup: int, ctx: Optional[Context] = None):
    """Basis.jl:37-53 -> (states, idxmap); states come from the device unranking."""
    _validate_basis_args(L, nup)
    m = Model(L, nup, [], np.zeros(L), [], ctx)
    return m.states, _RankMap(m)


# -------------------------------------------------------------- SpinModel.jl

def build_model(L: int, nup=None, hopping=(), onsite_field=None, zz=(), ctx: Optional[Context] = None) -> Model:
    """SpinModel.jl:23-38."""
    _validate_basis_args(L, nup)
    if onsite_field is None:
        onsite_field = np.zeros(L)
    return Model(L, nup, list(hopping), onsite_field, list(zz), ctx)


def nn_hopping(L: int, J: float):
    """SpinModel.jl:40-42."""
    return [(i, i + 1, J) for i in range(1, L)]


def long_range_hopping(L: int, J: Callable):
    """SpinModel.jl:44-46."""
    return [(i, j, J(i, j)) for i in range(1, L + 1) for j in range(i + 1, L + 1)]


def XXZChain(L: int, Jxy=1.0, Jz=1.0, hz=0.0, nup=None, boundary="open", ctx: Optional[Context] = None) -> Model:
    """SpinModel.jl:63-90."""
    hopping = [(i, i + 1, float(Jxy) / 2) for i in range(1, L)]
    zz = [(i, i + 1, float(Jz)) for i in range(1, L)]
    if boundary == "periodic":
        if L > 2:
            hopping.append((L, 1, float(Jxy) / 2))
            zz.append((L, 1, float(Jz)))
    elif boundary != "open":
        raise ValueError("boundary must be :open or :periodic")
    return build_model(L, nup=nup, hopping=hopping, onsite_field=np.full(L, float(hz)), zz=zz, ctx=ctx)


def momenta(model: Model):
    """SpinModel.jl:97-99."""
    return 2 * np.pi * np.arange(model.L) / model.L


# ------------------------------------------------------------ Hamiltonian.jl

def bit_at(state: int, i: int) -> int:
    """Hamiltonian.jl:19-21."""
    return (int(state) >> i) & 1


def sz_value(bit: int) -> float:
    """Hamiltonian.jl:23-25."""
    return 0.5 if bit == 1 else -0.5


def flip_bits(state: int, i: int, j: int) -> int:
    """Hamiltonian.jl:27-29."""
    return int(state) ^ (1 << i) ^ (1 << j)


def _is_dev(x) -> bool:
    return isinstance(x, DeviceVector)


def _up(model: Model, x, dtype=None) -> DeviceVector:
    """Host vector -> device (Julia Vector semantics); DeviceVector passes through."""
    if _is_dev(x):
        if dtype is not None and x.dtype != np.dtype(dtype):
            return x.astype(dtype)
        return x
    a = np.asarray(x)
    if dtype is not None:
        a = a.astype(dtype, copy=False)
    elif a.dtype not in (np.float64, np.complex128):
        a = a.astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    return model.to_device(a)


def apply_H_(out, psi, model: Model):
    """apply_H!(out, psi, model)   Hamiltonian.jl:211-273.  Returns `out`."""
    if _is_dev(out) and _is_dev(psi):
        check(lib().sd_apply_H(model._h, out._h, psi._h))
        return out
    out_a, psi_a = np.asarray(out), np.asarray(psi)
    if not (isinstance(out, np.ndarray) and out.flags.c_contiguous and out.flags.writeable):
        raise TypeError("out must be a writable contiguous numpy array or a DeviceVector")
    if out_a.shape != psi_a.shape:                                  # @assert length(out) == N
        raise ValueError("AssertionError: length(out) == N")
    if out_a.dtype != psi_a.dtype:                                   # `where T` dispatch
        raise TypeError("MethodError: out and psi must share one element type")
    if psi_a.shape != (model.dim,):
        raise ValueError("DimensionMismatch: psi does not match the model basis")
    psi_c = np.ascontiguousarray(psi_a)
    from .core import _sd_dtype
    check(lib().sd_apply_H_host(model._h, _sd_dtype(psi_c.dtype), _ptr(out_a), _ptr(psi_c)))
    return out


def apply_H_neg_(out, psi, model: Model):
    """The `-H` wrapper estimate_energy_bounds builds (Lanczos.jl:261-265)."""
    apply_H_(out, psi, model)
    if _is_dev(out):
        out.scale(-1.0)
    else:
        np.negative(out, out=out)
    return out


def apply_rescaled_H_(out, psi, applyH_, model: Model, a: float, b: float):
    """apply_rescaled_H!   Hamiltonian.jl:286-301: out = (H psi - b psi)/a."""
    _require_builtin(applyH_)
    if _is_dev(out) and _is_dev(psi):
        check(lib().sd_apply_rescaled_H(model._h, out._h, psi._h, float(a), float(b)))
        return out
    if len(out) != len(psi):
        raise ValueError("AssertionError: length(out) == length(psi)")
    d_psi = _up(model, psi)
    d_out = model.vector(d_psi.dtype)
    check(lib().sd_apply_rescaled_H(model._h, d_out._h, d_psi._h, float(a), float(b)))
    d_out.to_host(out)
    return out


def Sz_q_vector(model: Model, psi0, q: float, device: bool = False):
    """Sz_q_vector   Hamiltonian.jl:307-337."""
    n = model.dim if not _is_dev(psi0) else None
    if n is not None and len(psi0) != n:
        raise ValueError("AssertionError: length(psi0) == N")
    d_psi = _up(model, psi0)
    phi = model.vector(np.complex128)
    check(lib().sd_szq(model._h, phi._h, d_psi._h, float(q), None))
    return phi if (device or _is_dev(psi0)) else phi.to_host()


def _require_builtin(applyH_):
    if applyH_ not in (apply_H_, apply_H_neg_):
        raise NotImplementedError(
            "the GPU recurrences run the built-in matrix-free Hamiltonian; pass apply_H_ "
            "(the reference's public layer hard-codes Hamiltonian.apply_H! the same way, PublicAPI.jl:28,62,70,80)")


# ---------------------------------------------------------- InitialStates.jl

def _one_hot(model: Model, s: int, what: str, device: bool = False):
    idx = int(model.rank_of([s])[0])            # get(idxmap, s, 0) / Int(s)+1
    if idx == 0:
        raise ValueError(f"{what} is not contained in the model basis")
    if device:
        return model.vector(np.float64).set_onehot(idx - 1)
    psi0 = np.zeros(model.dim)
    psi0[idx - 1] = 1.0
    return psi0


def domain_wall_state(model: Model, device: bool = False):
    """InitialStates.jl:9-34."""
    nup = model.nup if model.mode == "sector" else int(np.ceil(model.L / 2))
    s = 0
    for i in range(nup):
        s |= 1 << i
    return _one_hot(model, s, "domain-wall state", device)


def neel_state(model: Model, device: bool = False):
    """InitialStates.jl:40-63."""
    s = 0
    for i in range(model.L):
        if (i + 1) % 2 == 1:
            s |= 1 << i
    return _one_hot(model, s, "Neel state", device)


def polarized_state(model: Model, up: bool = True, device: bool = False):
    """InitialStates.jl:70-90."""
    s = (1 << model.L) - 1 if up else 0
    return _one_hot(model, s, "requested polarized state", device)


def polarized_state_with_flips(model: Model, flips: Sequence[int], device: bool = False):
    """InitialStates.jl:98-130."""
    for site in flips:
        if not 1 <= site <= model.L:
            raise ValueError(f"flip site {site} is outside the model with L={model.L}")
    s = (1 << model.L) - 1
    for site in flips:
        s ^= 1 << (site - 1)
    return _one_hot(model, s, "requested flipped polarized state", device)


# ------------------------------------------------------------ Observables.jl

def _observables(psi, model: Model):
    d = _up(model, psi)
    mags, zz = np.zeros(model.L), np.zeros(model.L)
    check(lib().sd_vec_observables(d._h, _ptr(mags), _ptr(zz)))
    return mags, zz


def magnetization_per_site(psi, model: Model) -> np.ndarray:
    """magnetization_per_site(psi, model)   Observables.jl:14-37; psi may be a DeviceVector (nothing is downloaded)."""
    return _observables(psi, model)[0]


def connected_correlations(psi, model: Model) -> np.ndarray:
    """connected_correlations(psi, model)   Observables.jl:43-95:
    C_r = (1/L) sum_i <S_i S_{mod1(i+r,L)}> - <S_i><S_{mod1(i+r,L)}>, r = 0..L-1."""
    mags, zz = _observables(psi, model)
    L = model.L
    return np.array([(zz[r] - float(np.dot(mags, np.roll(mags, -r)))) / L for r in range(L)])


def structure_factor_Sq(psi, model: Model) -> dict:
    """structure_factor_Sq(psi, model)   Observables.jl:101-109: {q_n: Re fft(C_r)[n]}, q_n = 2 pi n / L."""
    S_q = np.fft.fft(connected_correlations(psi, model))
    return {2 * np.pi * n / model.L: float(S_q[n].real) for n in range(model.L)}


# ---------------------------------------------------------------- Lanczos.jl

def _eigvals_symtri(alpha, beta):
    from scipy.linalg import eigh_tridiagonal
    alpha = np.asarray(alpha, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    if len(alpha) == 1:
        return alpha.copy()
    return eigh_tridiagonal(alpha, beta, eigvals_only=True)


def _eigen_symtri(alpha, beta):
    from scipy.linalg import eigh_tridiagonal
    alpha = np.asarray(alpha, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    if len(alpha) == 1:
        return alpha.copy(), np.ones((1, 1))
    return eigh_tridiagonal(alpha, beta)


def randn_complex(rng: np.random.Generator, N: int) -> np.ndarray:
    """randn(rng, ComplexF64, N): real and imaginary parts N(0, 1/2)."""
    return (rng.standard_normal(N) + 1j * rng.standard_normal(N)) / np.sqrt(2.0)


def _start_vector(model: Model, v0, rng, cplx: bool) -> DeviceVector:
    """Start vectors stay a host responsibility (Lanczos.jl:39,99 use randn);
    a DeviceVector or a seed-filled vector avoids the host round trip at large N."""
    dt = np.complex128 if cplx else np.float64
    if v0 is not None:
        return _up(model, v0, dt)
    if model.ctx.world > 1:
        raise ValueError("multi-rank runs need an explicit DeviceVector start vector")
    rng = rng or np.random.default_rng()
    host = randn_complex(rng, model.dim) if cplx else rng.standard_normal(model.dim)
    return model.to_device(host)


def lanczos_extremal(applyH_, model: Model, lanc_m: int = 100, tol: float = 1e-12,
                     rng: Optional[np.random.Generator] = None, v0=None):
    """Lanczos.jl:27-84 -> (Emin, Emax) of the Krylov tridiagonal."""
    _require_builtin(applyH_)
    m = min(int(lanc_m), model.dim)
    d0 = _start_vector(model, v0, rng, True)
    alpha = np.zeros(m)
    beta = np.zeros(max(m - 1, 1))
    meff = ctypes.c_int()
    check(lib().sd_lanczos_extremal(model._h, d0._h, int(lanc_m), float(tol),
                                    1 if applyH_ is apply_H_neg_ else 0,
                                    _ptr(alpha), _ptr(beta), ctypes.byref(meff)))
    k = meff.value
    evals = _eigvals_symtri(alpha[:k], beta[:k - 1])
    return float(evals.min()), float(evals.max())


def lanczos_groundstate(applyH_, model: Model, lanc_m: int = 100, tol: float = 1e-12,
                        orthogonalize_tol: float = 1e-10, rng: Optional[np.random.Generator] = None,
                        v0=None, device: bool = False, return_tridiag: bool = False):
    """Lanczos.jl:87-181 -> (Emin, psi_gs).  psi_gs is a numpy array, or a
    DeviceVector when device=True (needed once N no longer fits the host)."""
    _require_builtin(applyH_)
    m = min(int(lanc_m), model.dim)
    d0 = _start_vector(model, v0, rng, False)
    alpha = np.zeros(m)
    beta = np.zeros(max(m - 1, 1))
    mact = ctypes.c_int()
    hV = ctypes.c_void_p()
    check(lib().sd_lanczos_groundstate(model._h, d0._h, int(lanc_m), float(tol), float(orthogonalize_tol),
                                       _ptr(alpha), _ptr(beta), ctypes.byref(mact), ctypes.byref(hV)))
    V = VecSet(model, hV, np.float64)
    ma = mact.value
    a_act = alpha[:ma]
    b_act = beta[:min(ma - 1, m - 1)]
    evals, evecs = _eigen_symtri(a_act, b_act)                      # :164-165
    idx = int(np.argmin(evals))
    Emin = float(evals[idx])
    y = evecs[:, idx]
    psi = model.vector(np.float64)
    n2 = V.lincomb(y.astype(np.complex128), psi)                    # :170
    psi.scale(1.0 / np.sqrt(n2))                                    # :171
    V.free()
    res = psi if device else psi.to_host()
    if return_tridiag:
        return Emin, res, a_act.copy(), b_act.copy()
    return Emin, res


def lanczos_groundstate_lean(applyH_, model: Model, lanc_m: int = 100, tol: float = 1e-12,
                             rng: Optional[np.random.Generator] = None, v0=None, device: bool = False,
                             return_tridiag: bool = False):
    """Memory-lean ground state -- an EXTENSION, not a reference function.  lanczos_groundstate
    keeps the N x m basis (Lanczos.jl:104: 481 GB at L = 32, m = 100); this runs the plain three-term recurrence on
    three device vectors twice from the same start vector: pass 1 for (alpha, beta), pass 2 to accumulate the Ritz
    vector V*y (sd_lanczos_lean).  Same return convention as lanczos_groundstate: (Emin, psi_gs), ||psi_gs|| = 1."""
    _require_builtin(applyH_)
    m = min(int(lanc_m), model.dim)
    d0 = _start_vector(model, v0, rng, False)
    alpha, beta = np.zeros(m), np.zeros(max(m, 1))
    meff = ctypes.c_int()
    check(lib().sd_lanczos_lean(model._h, d0._h, m, float(tol), _ptr(alpha), _ptr(beta), ctypes.byref(meff), None, None, None))
    k = meff.value
    theta, Q = _eigen_symtri(alpha[:k], beta[:k - 1])
    y = np.ascontiguousarray(Q[:, 0], dtype=np.float64)
    psi = model.vector(np.float64)
    n2 = ctypes.c_double()
    check(lib().sd_lanczos_lean(model._h, d0._h, k, float(tol), _ptr(alpha), _ptr(beta), ctypes.byref(meff), _ptr(y),
                                psi._h, ctypes.byref(n2)))
    psi.scale(1.0 / np.sqrt(n2.value))
    res = psi if (device or _is_dev(v0)) else psi.to_host()
    return (float(theta[0]), res, alpha[:k].copy(), beta[:k - 1].copy()) if return_tridiag else (float(theta[0]), res)


def lanczos_tridiag(applyH_, model: Model, v, lanc_m: int = 100, tol: float = 1e-12):
    """Lanczos.jl:196-246 -> (alpha, beta, normv)."""
    _require_builtin(applyH_)
    dv = _up(model, v, np.complex128)
    m = min(int(lanc_m), model.dim)
    alpha = np.zeros(m)
    beta = np.zeros(max(m - 1, 1))
    meff = ctypes.c_int()
    nv = ctypes.c_double()
    try:
        check(lib().sd_lanczos_tridiag(model._h, dv._h, int(lanc_m), float(tol), _ptr(alpha), _ptr(beta),
                                       ctypes.byref(meff), ctypes.byref(nv)))
    except Exception as e:
        from ._lib import ZeroNormError
        if isinstance(e, ZeroNormError):
            raise RuntimeError("starting vector has zero norm") from None
        raise
    k = meff.value
    return alpha[:k].copy(), beta[:k - 1].copy(), float(nv.value)


def estimate_energy_bounds(applyH_, model: Model, lanc_m: int = 80, rng: Optional[np.random.Generator] = None,
                           v0=None):
    """Lanczos.jl:255-271: two extremal runs, on H and on -H."""
    _require_builtin(applyH_)
    _, Emax = lanczos_extremal(apply_H_, model, lanc_m=lanc_m, rng=rng, v0=v0)
    _, Emax_neg = lanczos_extremal(apply_H_neg_, model, lanc_m=lanc_m, rng=rng, v0=v0)
    return -Emax_neg, Emax


# ------------------------------------------------------------- LanczosSqw.jl

def spectral_from_tridiagonal(alpha, beta, norm_phi, E0, w_range, eta=0.05, broaden="lorentz"):
    """LanczosSqw.jl:18-43 (host: W x m dense work)."""
    theta, Q = _eigen_symtri(alpha, beta)
    wts = np.abs(Q[0, :]) ** 2 * norm_phi ** 2
    w_range = np.asarray(w_range, dtype=np.float64)
    shifted = w_range[:, None] - (theta - E0)[None, :]
    if broaden == "lorentz":
        return ((1 / np.pi) * (eta / (shifted ** 2 + eta ** 2))) @ wts
    elif broaden == "gauss":
        return ((1 / (np.sqrt(2 * np.pi) * eta)) * np.exp(-(shifted ** 2) / (2 * eta ** 2))) @ wts
    raise RuntimeError(f"unknown broadening: {broaden}")


def _q_parallel(fn, psi0, model: Model, q_list, q_threads: int, **kw):
    """The reference runs the q-loop of lanczos_sqw / kpm_sqw under Threads.@threads (LanczosSqw.jl:65,
    KPM_Sqw.jl:218).  Same structure here: `q_threads` host threads, each with its OWN context (stream + scratch) and
    its own copy of the model and of psi0 on the same GPU, take a contiguous share of q_list, so the small kernels of
    different momenta overlap on the device (at L = 16 one apply fills 2 of 148 SMs).  Every q is computed by exactly
    the kernels of the sequential loop, so the result is bit-identical to q_threads = 1.  Single-GPU contexts only."""
    from concurrent.futures import ThreadPoolExecutor
    if model.ctx.world != 1:
        raise ValueError("q_threads > 1 needs a single-GPU context")
    host = psi0.to_host() if _is_dev(psi0) else np.asarray(psi0)
    q_list = list(q_list)
    nt = max(1, min(int(q_threads), len(q_list)))
    shares = [q_list[len(q_list) * t // nt: len(q_list) * (t + 1) // nt] for t in range(nt)]

    def work(qs):
        ctx = Context(model.ctx.device)
        m = None
        try:
            m = Model(model.L, model.nup, model.hopping_list, model.onsite_field, model.zz_list, ctx=ctx)
            return fn(host, m, qs, q_threads=1, **kw)
        finally:
            del m                                                    # vectors of fn are gone; model before its context
            ctx.close()

    with ThreadPoolExecutor(max_workers=nt) as ex:
        parts = list(ex.map(work, shares))
    return np.concatenate(parts, axis=0)


def _q_batches(model: Model, nq: int):
    """Momenta per call of the q-batched entry points: at most 128, and three [state][q] multi-vectors (16 B per state
    and padded column each) must fit in 80 % of the free device memory.  Returns [] when not even two columns fit."""
    free = ctypes.c_uint64()
    total = ctypes.c_uint64()
    check(lib().sd_ctx_mem_info(model.ctx._h, ctypes.byref(free), ctypes.byref(total)))
    fit = int(0.8 * free.value // (3 * 16 * max(model.dim, 1)))
    sizes = [n for n in (2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128) if n <= fit]     # the padded column counts
    if not sizes:
        return []
    cap = sizes[-1]
    out, i = [], 0
    while i < nq:
        out.append((i, min(nq, i + cap)))
        i += cap
    return out


# The multi-vector kernel is the one-thread-group-per-state gather: it wins where a single vector cannot fill the GPU
# (config 1, L = 16: 16 momenta in 2 launches per step instead of 48) and loses to the block kernel once that one is
# bandwidth-bound (L = 28: 3.7 ms per moment and momentum batched against 1.7 ms looped, profiles/round2_q_bench.json), so
# the automatic choice batches small bases only; q_batch=True / False overrides it.
Q_BATCH_AUTO_MAX_DIM = 1 << 20


def _use_q_batch(model: Model, q_list, q_threads, q_batch):
    if q_batch is None:
        q_batch = model.ctx.world == 1 and len(q_list) >= 2 and q_threads <= 1 and model.dim <= Q_BATCH_AUTO_MAX_DIM
    if q_batch and model.ctx.world != 1:
        raise ValueError("q_batch needs a single-GPU context")
    return bool(q_batch)


def lanczos_sqw(psi0, model: Model, q_list, w_range, lanc_m=200, eta=0.05, broaden="lorentz", q_threads: int = 1,
                q_batch: Optional[bool] = None):
    """LanczosSqw.jl:49-80.  q_batch (default on a single GPU with >= 2 momenta and at most Q_BATCH_AUTO_MAX_DIM states): all momenta as one interleaved
    [state][q] multi-vector, two fused kernels per Lanczos step for ALL of them (sd_lanczos_tridiag_szq_batch) -- the
    reference's Threads.@threads q-loop as data parallelism.  q_batch=False: one momentum after the other; q_threads > 1:
    that loop on several host threads / contexts (see _q_parallel)."""
    if q_threads > 1 and len(q_list) > 1 and not q_batch:
        return _q_parallel(lanczos_sqw, psi0, model, q_list, q_threads, w_range=w_range, lanc_m=lanc_m, eta=eta, broaden=broaden,
                           q_batch=False)
    psi0c = _up(model, psi0, np.complex128)
    tmp = model.vector(np.complex128)
    check(lib().sd_apply_H(model._h, tmp._h, psi0c._h))
    E0 = psi0c.dotu(tmp).real                                      # :59 dot(conj(psi0c), tmp)
    S = np.zeros((len(q_list), len(w_range)))
    if _use_q_batch(model, q_list, q_threads, q_batch):
        del tmp
        batches = _q_batches(model, len(q_list))
        if batches:
            qs = np.ascontiguousarray(q_list, dtype=np.float64)
            lm = int(lanc_m)
            for lo, hi in batches:
                n = hi - lo
                qb = qs[lo:hi].copy()
                alpha = np.zeros((n, lm))
                beta = np.zeros((n, lm))
                meff = np.zeros(n, dtype=np.int32)
                nphi = np.zeros(n)
                check(lib().sd_lanczos_tridiag_szq_batch(model._h, psi0c._h, _ptr(qb), n, lm, 1e-12,
                                                         _ptr(alpha), _ptr(beta), _ptr(meff), _ptr(nphi)))
                for c in range(n):
                    k = int(meff[c])
                    if k == 0:                                      # :69-72 norm(phi) == 0
                        continue
                    S[lo + c, :] = spectral_from_tridiagonal(alpha[c, :k], beta[c, :k - 1], float(nphi[c]), E0, w_range,
                                                             eta=eta, broaden=broaden)
            return S
    phi = model.vector(np.complex128)
    for iq, q in enumerate(q_list):                                 # :65 (the reference threads over q)
        n2 = ctypes.c_double()
        check(lib().sd_szq(model._h, phi._h, psi0c._h, float(q), ctypes.byref(n2)))
        if n2.value == 0.0:                                         # :69-72
            continue
        a, b, nphi = lanczos_tridiag(apply_H_, model, phi, lanc_m=lanc_m)
        S[iq, :] = spectral_from_tridiagonal(a, b, nphi, E0, w_range, eta=eta, broaden=broaden)
    return S


# ---------------------------------------------------------------- KPM_Sqw.jl

def _rescaling_from_bounds(E_min, E_max):
    """KPM_Sqw.jl:13-17."""
    return float((E_max - E_min) / (2 * 0.99)), float((E_max + E_min) / 2)


def get_rescaling_params(applyH_, model: Model, lanc_m=80, rng=None):
    """KPM_Sqw.jl:25-28."""
    E_min, E_max = estimate_energy_bounds(applyH_, model, lanc_m=lanc_m, rng=rng)
    return _rescaling_from_bounds(E_min, E_max)


def compute_chebyshev_moments(applyH_, phi, M: int, a: float, b: float, model: Model):
    """KPM_Sqw.jl:95-128: the fused moment loop (one kernel per moment)."""
    _require_builtin(applyH_)
    dphi = _up(model, phi, np.complex128)
    mu = np.zeros(int(M))
    check(lib().sd_kpm_moments(model._h, dphi._h, int(M), float(a), float(b), _ptr(mu)))
    return mu


def get_kernel(M: int, kernel: str):
    """KPM_Sqw.jl:131-145."""
    g = np.ones(M)
    if kernel == "jackson":
        n = np.arange(M)
        g = ((M - n + 1) * np.cos(np.pi * n / (M + 1))
             + np.sin(np.pi * n / (M + 1)) / np.tan(np.pi / (M + 1))) / (M + 1)
    elif kernel == "lorentz":
        lam = 3.0
        n = np.arange(M)
        g = np.sinh(lam * (1 - n / M)) / np.sinh(lam)
    return g


def kpm_sw(phi, applyH_, model: Model, w_range, a, b, E0, kpm_m=200, kernel="jackson"):
    """KPM_Sqw.jl:34-93."""
    mu = compute_chebyshev_moments(applyH_, phi, kpm_m, a, b, model)
    return _kpm_reconstruct(mu, w_range, a, b, E0, kpm_m, kernel)


def _kpm_reconstruct(mu, w_range, a, b, E0, kpm_m, kernel):
    """KPM_Sqw.jl:48-92: kernel-damped Chebyshev series on the frequency grid."""
    mu = mu * get_kernel(kpm_m, kernel)
    w_range = np.asarray(w_range, dtype=np.float64)
    x = (w_range + E0 - b) / a
    S = np.zeros(len(w_range))
    inside = np.abs(x) < 1.0
    xi = x[inside]
    Tm2 = np.ones_like(xi)
    acc = mu[0] * Tm2
    if kpm_m >= 2:
        Tm1 = xi.copy()
        acc = acc + 2.0 * mu[1] * Tm1
        for n in range(2, kpm_m):
            Tn = 2.0 * xi * Tm1 - Tm2
            acc = acc + 2.0 * mu[n] * Tn
            Tm2, Tm1 = Tm1, Tn
    S[inside] = np.maximum(0.0, acc / (a * np.pi * np.sqrt(1.0 - xi ** 2)))
    return S


def kpm_sqw(psi0, model: Model, q_list, w_range, a=None, b=None, kpm_m=200, kernel="jackson", rng=None, q_threads: int = 1,
            q_batch: Optional[bool] = None):
    """KPM_Sqw.jl:191-256.  q_batch (default on a single GPU with >= 2 momenta and at most Q_BATCH_AUTO_MAX_DIM states): the moments of all momenta from one
    interleaved [state][q] multi-vector, one fused kernel per moment for ALL of them (sd_kpm_moments_szq_batch);
    q_batch=False: one momentum after the other; q_threads > 1: that loop on several host threads / contexts."""
    if q_threads > 1 and len(q_list) > 1 and not q_batch:
        if a is None or b is None:                                  # :211-215 once, not per thread
            a, b = get_rescaling_params(apply_H_, model, rng=rng)
        return _q_parallel(kpm_sqw, psi0, model, q_list, q_threads, w_range=w_range, a=a, b=b, kpm_m=kpm_m, kernel=kernel,
                           q_batch=False)
    if _use_q_batch(model, q_list, q_threads, q_batch):
        psi0c = _up(model, psi0, np.complex128)
        tmp = model.vector(np.complex128)
        r = SdComplex()
        check(lib().sd_apply_H_dot(model._h, tmp._h, psi0c._h, ctypes.byref(r)))    # :207-209 fused E0
        E0 = float(r.re)
        del tmp
        if a is None or b is None:
            a, b = get_rescaling_params(apply_H_, model, rng=rng)
        batches = _q_batches(model, len(q_list))
        S = np.zeros((len(q_list), len(w_range)))
        qs = np.ascontiguousarray(q_list, dtype=np.float64)
        ok = bool(batches)
        for lo, hi in batches:
            n = hi - lo
            qb = qs[lo:hi].copy()
            mu = np.zeros((n, int(kpm_m)))
            nphi = np.zeros(n)
            blown = ctypes.c_int()
            check(lib().sd_kpm_moments_szq_batch(model._h, psi0c._h, _ptr(qb), n, int(kpm_m), float(a), float(b),
                                                 _ptr(mu), _ptr(nphi), ctypes.byref(blown)))
            if blown.value:                                         # :117-121 renormalisation needed: per-momentum path
                ok = False
                break
            for c in range(n):
                if nphi[c] == 0:                                    # :226-229
                    continue
                S[lo + c, :] = nphi[c] ** 2 * _kpm_reconstruct(mu[c], w_range, a, b, E0, int(kpm_m), kernel)
        if ok:
            return S
        del psi0c
    psi0c = _up(model, psi0, np.complex128)
    S = np.zeros((len(q_list), len(w_range)))
    tmp = model.vector(np.complex128)
    r = SdComplex()
    check(lib().sd_apply_H_dot(model._h, tmp._h, psi0c._h, ctypes.byref(r)))    # :207-209 fused E0
    E0 = float(r.re)
    if a is None or b is None:
        a, b = get_rescaling_params(apply_H_, model, rng=rng)
    phi = model.vector(np.complex128)
    for iq, q in enumerate(q_list):
        n2 = ctypes.c_double()
        check(lib().sd_szq(model._h, phi._h, psi0c._h, float(q), ctypes.byref(n2)))
        norm_phi = float(np.sqrt(n2.value))
        if norm_phi == 0:
            continue
        phi.scale(1.0 / norm_phi)                                   # :231
        Sq = kpm_sw(phi, apply_H_, model, w_range, a=a, b=b, E0=E0, kpm_m=kpm_m, kernel=kernel)
        S[iq, :] = norm_phi ** 2 * Sq
    return S


# ---------------------------------------------------- TimeEvolution/Krylov.jl

def krylov_time_evolve(psi0, dt: float, applyH_, model: Model, kry_m: int = 30, device: bool = False):
    """Krylov.jl:136-192."""
    _require_builtin(applyH_)
    d0 = _up(model, psi0)
    n = int(kry_m)
    alpha = np.zeros(n, dtype=np.complex128)
    beta = np.zeros(max(n - 1, 1))
    meff = ctypes.c_int()
    norm0 = ctypes.c_double()
    hV = ctypes.c_void_p()
    check(lib().sd_krylov_basis(model._h, d0._h, n, _ptr(alpha), _ptr(beta), ctypes.byref(meff),
                                ctypes.byref(norm0), ctypes.byref(hV)))
    V = VecSet(model, hV, d0.dtype)
    if norm0.value == 0.0:                                          # :148
        V.free()
        return d0 if _is_dev(psi0) else np.array(psi0, copy=True)
    k = meff.value
    al, be = alpha[:k], beta[:k - 1].astype(np.complex128)
    TR = np.diag(al) + np.diag(be, 1) + np.diag(be, -1)             # :175
    if np.array_equal(TR, TR.conj().T):
        D, Q = np.linalg.eigh(TR)
    else:
        D, Q = np.linalg.eig(TR)
    U_T = Q @ np.diag(np.exp(-1j * D * dt)) @ Q.conj().T            # :180
    e1 = np.zeros(k, dtype=np.complex128)
    e1[0] = norm0.value
    y = U_T @ e1
    psit = model.vector(np.complex128)
    n2 = V.lincomb(y, psit)                                         # :185-188
    psit.scale(1.0 / np.sqrt(n2))                                   # :190
    V.free()
    return psit if (device or _is_dev(psi0)) else psit.to_host()


class KrylovWorkspace:
    """Krylov.jl:25-40.  The reference preallocates m basis vectors + w on the host; on the GPU path the basis is
    device-resident and owned by the library for the duration of one call (sd_krylov_basis -> sd_vecset), so the
    workspace only carries the sizes the reference asserts on (`length(ws.V) >= kry_m`)."""

    def __init__(self, n: int, m: int):
        self.n, self.m = int(n), int(m)
        self.alpha = np.zeros(self.m, dtype=np.complex128)
        self.beta = np.zeros(max(self.m - 1, 0), dtype=np.complex128)


def krylov_time_evolve_(psi_out, psi_in, dt: float, applyH_, model: Model, ws: KrylovWorkspace, kry_m: int = 30):
    """krylov_time_evolve!(psi_out, psi_in, dt, applyH!, model, ws; kry_m)   Krylov.jl:55-118: the in-place twin
    (untested upstream), routed to the same device routine as krylov_time_evolve.  Returns None; psi_out is
    returned when psi_in has zero norm, as upstream does (:69-72)."""
    cplx = (lambda x: (x.dtype if _is_dev(x) else np.asarray(x).dtype) == np.complex128)
    if not (cplx(psi_out) and cplx(psi_in)):
        raise TypeError("MethodError: krylov_time_evolve! takes Vector{ComplexF64}")
    if len(psi_out) != len(psi_in):
        raise ValueError("AssertionError: length(psi_out) == n")
    if ws.m < int(kry_m):
        raise ValueError("AssertionError: length(ws.V) >= kry_m")
    res = krylov_time_evolve(psi_in, float(dt), applyH_, model, kry_m=int(kry_m), device=_is_dev(psi_out))
    zero = res is psi_in or (not _is_dev(psi_in) and not _is_dev(res) and not np.any(res))
    if _is_dev(psi_out):
        if res is not psi_out:
            psi_out.copy_from(res if _is_dev(res) else model.to_device(np.asarray(res, dtype=np.complex128)))
    else:
        psi_out[:] = res.to_host() if _is_dev(res) else res
    return psi_out if zero else None


# ------------------------------------------------- TimeEvolution/Chebyshev.jl

class ChebyshevWorkspace:
    """Chebyshev.jl:19-36.  phi_prev / phi_curr / phi_next / psi_t live on the device inside
    sd_chebyshev_evolve (`w` is unused upstream as well); the workspace keeps N for the size check (:86)."""

    def __init__(self, N_or_psi, dtype=np.complex128):
        self.N = int(N_or_psi) if np.isscalar(N_or_psi) else len(N_or_psi)
        self.dtype = np.dtype(dtype)


_MINUS_I_POW = (1.0 + 0j, -1j, -1.0 + 0j, 1j)


def chebyshev_coefficients(dt, cheb_n, Ebounds):
    """Chebyshev.jl:70-79 (host: Bessel functions)."""
    from scipy.special import jv
    E_min, E_max = Ebounds
    a = (E_max - E_min) / (2 * 0.9999)
    b = (E_max + E_min) / 2
    phase_factor = np.exp(-1j * b * dt)
    c = np.empty(cheb_n, dtype=np.complex128)
    for k in range(cheb_n):
        c[k] = (2 - (1.0 if k == 0 else 0.0)) * _MINUS_I_POW[k % 4] * jv(k, a * dt) * phase_factor
    return c, a, b


def chebyshev_time_evolve(psi0, dt: float, applyH_, model: Model, cheb_n: int = 100, Ebounds=(-1.0, 1.0),
                          device: bool = False, workspace: Optional[ChebyshevWorkspace] = None):
    """Chebyshev.jl:61-124."""
    _require_builtin(applyH_)
    if not cheb_n >= 1:
        raise ValueError("AssertionError: cheb_n must be >= 1")
    if workspace is not None and workspace.N != len(psi0):
        raise ValueError("AssertionError: Workspace size mismatch")             # :86
    dt_in = psi0.dtype if _is_dev(psi0) else np.asarray(psi0).dtype
    if dt_in != np.complex128:
        raise TypeError("InexactError: real psi0 cannot hold complex Chebyshev sums")
    c, a, b = chebyshev_coefficients(dt, int(cheb_n), Ebounds)
    d0 = _up(model, psi0, np.complex128)
    out = model.vector(np.complex128)
    check(lib().sd_chebyshev_evolve(model._h, d0._h, _ptr(c), int(cheb_n), float(a), float(b), out._h))
    return out if (device or _is_dev(psi0)) else out.to_host()


# ------------------------------------------------- TimeEvolution/KPM.jl (site-resolved KPM)

def site_sz_operator(i: int):
    """The operator `create_spin_operator(i, :z)` that kpm_correlation_matrix builds (TimeEvolution/KPM.jl:205-206),
    restricted to what the hot path covers: S^z on 1-based site i, psi -> S^z_i psi (complex result, like every other
    operator output of the reference).  Device in, device out; host in, host out."""
    def op(psi, model: Model):
        if not 1 <= i <= model.L:
            raise ValueError("site index out of range")
        w = (SdComplex * model.L)()
        w[i - 1].re = 1.0
        d = _up(model, psi)
        out = model.vector(np.complex128)
        check(lib().sd_apply_sz_weights(model._h, out._h, d._h, ctypes.cast(w, ctypes.c_void_p), None))
        return out if _is_dev(psi) else out.to_host()
    return op


def kpm_get_rescaling_params(applyH_, model: Model, lanc_m: int = 80, rng=None):
    """TimeEvolution/KPM.jl:45-49 -- note the factor: a = (Emax - Emin) / 2 * 0.9 (NOT / 0.99 as in KPM_Sqw.jl:13-17)."""
    E_min, E_max = estimate_energy_bounds(applyH_, model, lanc_m=lanc_m, rng=rng)
    return float((E_max - E_min) / 2 * 0.9), float((E_max + E_min) / 2)


def get_jackson_kernel(n: int):
    """TimeEvolution/KPM.jl:170-177."""
    k = np.arange(n)
    d = np.pi / (n + 1)
    return ((n - k + 1) * np.cos(d * k) + np.sin(d * k) / np.tan(d)) / (n + 1)


def evaluate_chebyshev_series(mu, x: float, a: float) -> float:
    """TimeEvolution/KPM.jl:184-206: 0 outside (-1, 1); sum_k mu_k T_k(x) / (pi sqrt(1 - x^2)) * (2 / a) (no factor 2 on k >= 1)."""
    if abs(x) >= 1.0:
        return 0.0
    n = len(mu)
    total = mu[0]
    if n > 1:
        total += mu[1] * x
    Tp, Tc = 1.0, x
    for k in range(2, n):
        Tn = 2 * x * Tc - Tp
        total += mu[k] * Tn
        Tp, Tc = Tc, Tn
    return float(total / (np.pi * np.sqrt(1 - x * x)) * (2 / a))


def compute_cross_chebyshev_moments(chi, phi, n: int, a: float, b: float, applyH_, model: Model):
    """TimeEvolution/KPM.jl:121-165: mu_k = <chi| T_k(H~) |phi> with phi normalised first and `dot(conj(chi), .)`, i.e.
    the UNCONJUGATED product sum_i chi_i phi_i (sd_vec_dotu).  The recurrence runs on the device: one fused Chebyshev
    step per moment (sd_cheb_step: 2 (H v - b v) / a - v_prev in the apply kernel's epilogue) plus one dotu pass."""
    _require_builtin(applyH_)
    n = int(n)
    dchi = _up(model, chi, np.complex128)
    dphi = _up(model, phi, np.complex128)
    norm_phi = dphi.norm()
    prev = model.vector(np.complex128).copy_from(dphi)
    prev.scale(1.0 / norm_phi)
    curr = model.vector(np.complex128)
    check(lib().sd_apply_rescaled_H(model._h, curr._h, prev._h, float(a), float(b)))
    mom = np.zeros(n)

    def real_or_inexact(z):
        if abs(z.imag) > 0.0 and abs(z.imag) > 1e-300:
            raise TypeError("InexactError: complex moment assigned to a Float64 array")   # moments[1] = dot(...) :147-148
        return z.real

    mom[0] = real_or_inexact(dchi.dotu(prev) * norm_phi)
    if n > 1:
        mom[1] = real_or_inexact(dchi.dotu(curr) * norm_phi)
    zero = SdComplex(0.0, 0.0)
    for k in range(2, n):
        # phi_next = 2 H~ phi_curr - phi_prev, written over phi_prev (same element, read before it is written)
        check(lib().sd_cheb_step(model._h, prev._h, curr._h, prev._h, float(a), float(b), None, None, None, None, zero))
        prev, curr = curr, prev
        mom[k] = (dchi.dotu(curr)).real * norm_phi
    return mom


def kpm_dynamical_correlation(psi, operator_A, operator_B, w_range, applyH_, model: Model, n: int = 300, eps: float = 0.1,
                              a=None, b=None):
    """TimeEvolution/KPM.jl:74-118: S(w) = <psi| A^dag delta(w - H) B |psi> from n Jackson-damped cross moments; x = (w - b) / a
    (no E0 shift, as in the reference); negative values clipped to 0."""
    if a is None or b is None:
        a, b = kpm_get_rescaling_params(applyH_, model, lanc_m=n)
    phi = operator_B(psi, model)
    chi = operator_A(psi, model)
    mu = compute_cross_chebyshev_moments(chi, phi, n, a, b, applyH_, model) * get_jackson_kernel(n)
    S = np.array([evaluate_chebyshev_series(mu, (w - b) / a, a) for w in np.asarray(w_range, dtype=np.float64)])
    return np.maximum(S, 0.0)


def kpm_correlation_matrix(psi, w_range, applyH_, model: Model, n: int = 300, eps: float = 0.1):
    """TimeEvolution/KPM.jl:211-232 with opA = opB = S^z: C[i, j, :] = |S_ij(w)|, rescaling parameters computed once."""
    L = model.L
    w_range = np.asarray(w_range, dtype=np.float64)
    C = np.zeros((L, L, len(w_range)))
    a, b = kpm_get_rescaling_params(applyH_, model)
    dpsi = _up(model, psi)
    for i in range(1, L + 1):
        for j in range(1, L + 1):
            C[i - 1, j - 1, :] = np.abs(kpm_dynamical_correlation(dpsi, site_sz_operator(i), site_sz_operator(j), w_range,
                                                                 applyH_, model, n=n, eps=eps, a=a, b=b))
    return C


def Sqw(C, q: float, positions):
    """TimeEvolution/KPM.jl:236-246: S(q, w) = (1/N) sum_ij Re(e^{-iq (r_i - r_j)} C_ij(w))."""
    positions = np.asarray(positions, dtype=np.float64)
    N = len(positions)
    ph = np.exp(-1j * q * (positions[:, None] - positions[None, :]))
    return np.einsum("ij,ijw->w", ph.real, C) / N


# -------------------------------------------------------------- PublicAPI.jl

def groundstate(model: Model, method="lanczos", **kw):
    """PublicAPI.jl:25-35."""
    if method == "lanczos":
        return lanczos_groundstate(apply_H_, model, **kw)
    if method == "lanczos_lean":                       # extension: three work vectors instead of the N x m basis
        return lanczos_groundstate_lean(apply_H_, model, **kw)
    raise ValueError(f"unsupported ground-state method: {method}")


def time_evolve(model: Model, psi0, t, method="krylov", Ebounds=None, **kw):
    """PublicAPI.jl:50-88."""
    if method == "krylov":
        return krylov_time_evolve(psi0, float(t), apply_H_, model, **kw)
    elif method == "chebyshev":
        bounds = estimate_energy_bounds(apply_H_, model) if Ebounds is None else Ebounds
        return chebyshev_time_evolve(psi0, float(t), apply_H_, model, Ebounds=bounds, **kw)
    raise ValueError(f"unsupported time-evolution method: {method}")


def structure_factor(model: Model, psi):
    """structure_factor(model, psi)   PublicAPI.jl:94-106."""
    return structure_factor_Sq(psi, model)


def dynamical_structure_factor(model: Model, psi0, q, w, method="lanczos", **kw):
    """PublicAPI.jl:122-155."""
    q_list = np.asarray(q, dtype=np.float64)
    w_range = np.asarray(w, dtype=np.float64)
    if method == "lanczos":
        return lanczos_sqw(psi0, model, q_list, w_range, **kw)
    elif method == "kpm":
        return kpm_sqw(psi0, model, q_list, w_range, **kw)
    raise ValueError(f"unsupported dynamical structure-factor method: {method}")
