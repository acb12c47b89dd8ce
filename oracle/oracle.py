"""CPU oracle: numpy/C restatement of SpinDynamics.jl's H.psi path and the
recurrences that call it.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the
checker or the timed CPU baseline.  The product (spindynamics.jl_b200/) never
imports it.

The reference is Julia and cannot run in this image (no `julia`), so this file
follows the reference line by line instead; every function cites the
reference file:line (relative to /root/reference) it restates.  Host-side dense
work that the reference delegates to LinearAlgebra/SpecialFunctions uses
numpy/scipy here (`eigh_tridiagonal`, `eig`, `jv`).

Pins (tests/test_oracle.py): reference known answers (test_PublicAPI.jl:5-28,
:40-51, :56-118; test_Lanczos.jl:6-54; test_Hamiltonian.jl:93-110;
test_KPM.jl:67-91; test_Basis.jl:4-19; test_InitialStates.jl) and an
independent Kronecker-product construction of the XXZ Hamiltonian.  Ordering
inside a sector follows Combinatorics.jl's documented lexicographic order and
is NOT pinned by any reference test ("parity unpinned" for that property).

Julia `f!` names become `f_` here; 1-based indices stay 1-based wherever the
reference exposes them (idxmap values, bond sites).
"""
from __future__ import annotations

import ctypes
import itertools
import os
import subprocess
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Bond(ctypes.Structure):
    _fields_ = [("i", ctypes.c_int64), ("j", ctypes.c_int64), ("J", ctypes.c_double)]


def build_c(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc + OpenMP)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_c())
        u64, i64, dbl, vp = ctypes.c_uint64, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
        L.orc_model_create.restype = vp
        L.orc_model_create.argtypes = [ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp,
                                       ctypes.c_int, vp]
        L.orc_model_free.argtypes = [vp]
        L.orc_model_dim.restype = u64
        L.orc_model_dim.argtypes = [vp]
        L.orc_model_states.restype = ctypes.POINTER(u64)
        L.orc_model_states.argtypes = [vp]
        L.orc_rank.argtypes = [vp, vp, u64, vp]
        L.orc_apply_H_f64.argtypes = [vp, vp, vp]
        L.orc_apply_H_c128.argtypes = [vp, vp, vp]
        L.orc_apply_H_f64_ranked.argtypes = [vp, vp, vp]
        L.orc_apply_rescaled_H.argtypes = [vp, vp, vp, dbl, dbl, ctypes.c_int]
        L.orc_szq.argtypes = [vp, vp, vp, ctypes.c_int, dbl]
        L.orc_fill_seeded.argtypes = [vp, u64, u64, u64, ctypes.c_int]
        L.orc_build_sector_basis.argtypes = [ctypes.c_int, ctypes.c_int, vp]
        L.orc_rank_closed_form.restype = u64
        L.orc_rank_closed_form.argtypes = [ctypes.c_int, ctypes.c_int, u64]
        L.orc_row_seeded_f64.restype = dbl
        L.orc_row_seeded_f64.argtypes = [ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp,
                                         ctypes.c_int, vp, u64, u64, dbl]
        L.orc_num_threads.restype = ctypes.c_int
        L.orc_set_num_threads.argtypes = [ctypes.c_int]
        L.orc_set_num_threads.restype = None
        L.orc_sector_dim.restype = u64
        L.orc_sector_dim.argtypes = [ctypes.c_int, ctypes.c_int]
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------- Basis.jl

def _validate_basis_args(L: int, nup: Optional[int] = None) -> None:
    """Basis.jl:9-20 (ArgumentError -> ValueError)."""
    if not L >= 1:
        raise ValueError("L must be at least 1")
    if not L <= 63:
        raise ValueError("L must be at most 63 when using UInt64 basis states")
    if nup is not None and not (0 <= nup <= L):
        raise ValueError("nup must satisfy 0 <= nup <= L")


def build_full_basis(L: int):
    """Basis.jl:23-34."""
    _validate_basis_args(L)
    N = 1 << L
    states = np.arange(N, dtype=np.uint64)
    idxmap = {int(s): i + 1 for i, s in enumerate(states)}
    return states, idxmap


def build_sector_basis(L: int, nup: int):
    """Basis.jl:37-53: `for comb in combinations(1:L, nup)`; site i -> bit i-1.
    itertools.combinations yields the same lexicographic order."""
    _validate_basis_args(L, nup)
    states = []
    for comb in itertools.combinations(range(1, L + 1), nup):
        s = 0
        for i in comb:
            s |= 1 << (i - 1)
        states.append(s)
    states = np.array(states, dtype=np.uint64)
    idxmap = {int(s): i + 1 for i, s in enumerate(states)}
    return states, idxmap


# ------------------------------------------------------------- SpinModel.jl

@dataclass
class Model:
    """SpinModel.jl:6-15.  `states` comes from the C oracle for speed (it is
    checked against build_sector_basis above in tests); `idxmap` is built
    lazily because a Python dict of 10^7 entries is only needed by small tests."""
    L: int
    nup: Optional[int]
    mode: str                      # "full" | "sector"
    hopping_list: list
    onsite_field: np.ndarray
    zz_list: list
    _c: int = field(default=0, repr=False)
    _idxmap: Optional[dict] = field(default=None, repr=False)

    @property
    def states(self) -> np.ndarray:
        n = lib().orc_model_dim(self._c)
        p = lib().orc_model_states(self._c)
        return np.ctypeslib.as_array(p, shape=(n,))

    @property
    def idxmap(self) -> dict:
        if self._idxmap is None:
            self._idxmap = {int(s): i + 1 for i, s in enumerate(self.states)}
        return self._idxmap

    def __len__(self):
        return int(lib().orc_model_dim(self._c))

    def __del__(self):
        try:
            if self._c:
                lib().orc_model_free(self._c)
                self._c = 0
        except Exception:
            pass


def _bond_array(lst):
    arr = (_Bond * max(1, len(lst)))()
    for k, (i, j, J) in enumerate(lst):
        arr[k].i, arr[k].j, arr[k].J = int(i), int(j), float(J)
    return arr


def build_model(L: int, nup: Optional[int] = None, hopping=(), onsite_field=None, zz=()):
    """SpinModel.jl:23-38."""
    _validate_basis_args(L, nup)
    if onsite_field is None:
        onsite_field = np.zeros(L)
    hop = [(int(i), int(j), float(J)) for (i, j, J) in hopping]
    zzl = [(int(i), int(j), float(J)) for (i, j, J) in zz]
    fld = np.ascontiguousarray(onsite_field, dtype=np.float64)
    ha, za = _bond_array(hop), _bond_array(zzl)
    c = lib().orc_model_create(L, -1 if nup is None else nup,
                               ctypes.cast(ha, ctypes.c_void_p), len(hop),
                               ctypes.cast(za, ctypes.c_void_p), len(zzl), _ptr(fld))
    if not c:
        raise ValueError("invalid basis arguments")
    return Model(L, nup, "full" if nup is None else "sector", hop, fld, zzl, _c=c)


def nn_hopping(L: int, J: float):
    """SpinModel.jl:40-42."""
    return [(i, i + 1, J) for i in range(1, L)]


def long_range_hopping(L: int, J: Callable):
    """SpinModel.jl:44-46."""
    return [(i, j, J(i, j)) for i in range(1, L + 1) for j in range(i + 1, L + 1)]


def XXZChain(L: int, Jxy=1.0, Jz=1.0, hz=0.0, nup=None, boundary="open"):
    """SpinModel.jl:63-90: hop (i,i+1,Jxy/2), zz (i,i+1,Jz), field fill(hz,L);
    periodic closes the ring only if L > 2."""
    hopping = [(i, i + 1, float(Jxy) / 2) for i in range(1, L)]
    zz = [(i, i + 1, float(Jz)) for i in range(1, L)]
    if boundary == "periodic":
        if L > 2:
            hopping.append((L, 1, float(Jxy) / 2))
            zz.append((L, 1, float(Jz)))
    elif boundary != "open":
        raise ValueError("boundary must be :open or :periodic")
    return build_model(L, nup=nup, hopping=hopping, onsite_field=np.full(L, float(hz)), zz=zz)


def momenta(model: Model):
    """SpinModel.jl:97-99."""
    return 2 * np.pi * np.arange(model.L) / model.L


# ----------------------------------------------------------- Hamiltonian.jl

def apply_H_(out: np.ndarray, psi: np.ndarray, model: Model) -> np.ndarray:
    """Hamiltonian.jl:211-273 via the C restatement (oracle.c)."""
    assert out.shape == psi.shape and out.dtype == psi.dtype      # :220 + `where T`
    assert out.flags.c_contiguous and psi.flags.c_contiguous
    assert psi.shape[0] == len(model)
    if psi.dtype == np.float64:
        lib().orc_apply_H_f64(model._c, _ptr(out), _ptr(psi))
    elif psi.dtype == np.complex128:
        lib().orc_apply_H_c128(model._c, _ptr(out), _ptr(psi))
    else:
        raise TypeError(psi.dtype)
    return out


def apply_H_ranked_(out: np.ndarray, psi: np.ndarray, model: Model) -> np.ndarray:
    """NOT the reference's algorithm: apply_H! with the Dict probe replaced by combinatorial ranking
    (oracle.c orc_apply_H_f64_ranked) -- the second, stronger CPU baseline bench.py reports."""
    assert out.shape == psi.shape and psi.dtype == np.float64 and out.dtype == np.float64
    assert out.flags.c_contiguous and psi.flags.c_contiguous and psi.shape[0] == len(model)
    lib().orc_apply_H_f64_ranked(model._c, _ptr(out), _ptr(psi))
    return out


def apply_H_np_(out: np.ndarray, psi: np.ndarray, model: Model) -> np.ndarray:
    """Second, independent restatement of Hamiltonian.jl:211-273 in pure numpy
    (vectorised over idx, dict replaced by a sorted-array search).  Used to
    cross-check the C restatement."""
    st = np.array(model.states, dtype=np.uint64)
    L = model.L
    one = np.uint64(1)

    def sz(bitpos):
        return np.where((st >> np.uint64(bitpos)) & one == one, 0.5, -0.5)

    diag = np.zeros(len(st))
    for i in range(1, L + 1):
        diag = diag + model.onsite_field[i - 1] * sz(i - 1)
    for (i, j, Jz) in model.zz_list:
        diag = diag + Jz * sz(i - 1) * sz(j - 1)
    val = diag * psi
    if model.mode == "sector":
        order = np.argsort(st, kind="stable")
        sorted_states = st[order]
    for (i, j, Jxy) in model.hopping_list:
        bi = (st >> np.uint64(i - 1)) & one
        bj = (st >> np.uint64(j - 1)) & one
        act = bi != bj
        ns = st ^ (one << np.uint64(i - 1)) ^ (one << np.uint64(j - 1))
        if model.mode == "full":
            tgt = ns.astype(np.int64)
            ok = act
        else:
            pos = np.searchsorted(sorted_states, ns)
            pos = np.minimum(pos, len(st) - 1)
            ok = act & (sorted_states[pos] == ns)
            tgt = order[pos]
        contrib = np.where(ok, Jxy * psi[tgt], 0)
        val = val + contrib
    out[:] = val
    return out


def apply_rescaled_H_(out, psi, applyH_, model, a: float, b: float):
    """Hamiltonian.jl:286-301: out = (H psi - b psi)/a."""
    assert len(out) == len(psi)
    applyH_(out, psi, model)
    out[:] = (out - b * psi) / a
    return out


def Sz_q_vector(model: Model, psi0: np.ndarray, q: float) -> np.ndarray:
    """Hamiltonian.jl:307-337 (C restatement)."""
    N = len(model)
    assert psi0.shape[0] == N
    phi = np.zeros(N, dtype=np.complex128)
    p = np.ascontiguousarray(psi0)
    if p.dtype == np.float64:
        lib().orc_szq(model._c, _ptr(phi), _ptr(p), 0, float(q))
    else:
        p = p.astype(np.complex128, copy=False)
        lib().orc_szq(model._c, _ptr(phi), _ptr(p), 1, float(q))
    return phi


def Sz_q_vector_np(model: Model, psi0: np.ndarray, q: float) -> np.ndarray:
    """Independent numpy restatement of Hamiltonian.jl:307-337."""
    st = np.array(model.states, dtype=np.uint64)
    L = model.L
    phases = np.exp(1j * q * np.arange(L))
    sq = np.zeros(len(st), dtype=np.complex128)
    for r in range(L):
        sq = sq + phases[r] * np.where((st >> np.uint64(r)) & np.uint64(1) == 1, 0.5, -0.5)
    return (1.0 / np.sqrt(L)) * sq * psi0.astype(np.complex128)


# ---------------------------------------------------------- InitialStates.jl

def _one_hot(model: Model, s: int, what: str) -> np.ndarray:
    psi0 = np.zeros(len(model))
    if model.mode == "full":
        idx = s + 1
    else:
        r = np.zeros(1, dtype=np.int64)
        lib().orc_rank(model._c, _ptr(np.array([s], dtype=np.uint64)), 1, _ptr(r))
        idx = int(r[0])
    if idx == 0:
        raise ValueError(f"{what} is not contained in the model basis")
    psi0[idx - 1] = 1.0
    return psi0


def domain_wall_state(model: Model):
    """InitialStates.jl:9-34."""
    nup = model.nup if model.mode == "sector" else int(np.ceil(model.L / 2))
    s = 0
    for i in range(nup):
        s |= 1 << i
    return _one_hot(model, s, "domain-wall state")


def neel_state(model: Model):
    """InitialStates.jl:40-63: up on odd sites (1-based)."""
    s = 0
    for i in range(model.L):
        if (i + 1) % 2 == 1:
            s |= 1 << i
    return _one_hot(model, s, "Neel state")


def polarized_state(model: Model, up: bool = True):
    """InitialStates.jl:70-90."""
    s = (1 << model.L) - 1 if up else 0
    return _one_hot(model, s, "requested polarized state")


def polarized_state_with_flips(model: Model, flips: Sequence[int]):
    """InitialStates.jl:98-130."""
    for site in flips:
        if not 1 <= site <= model.L:
            raise ValueError(f"flip site {site} is outside the model with L={model.L}")
    s = (1 << model.L) - 1
    for site in flips:
        s ^= 1 << (site - 1)
    return _one_hot(model, s, "requested flipped polarized state")


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
    """Julia randn(rng, ComplexF64, N): real and imaginary parts N(0, 1/2).
    The stream itself is not reproducible outside Julia, so start vectors are
    inputs everywhere in this build."""
    return (rng.standard_normal(N) + 1j * rng.standard_normal(N)) / np.sqrt(2.0)


def lanczos_extremal(applyH_, model: Model, lanc_m: int = 100, tol: float = 1e-12,
                     rng: Optional[np.random.Generator] = None, v0: Optional[np.ndarray] = None):
    """Lanczos.jl:27-84.  `v0` replaces randn(rng, ComplexF64, N) when given."""
    N = len(model)
    m = min(lanc_m, N)
    if v0 is None:
        rng = rng or np.random.default_rng()
        psi0 = randn_complex(rng, N)
    else:
        psi0 = np.array(v0, dtype=np.complex128)
    psi0 = psi0 / np.linalg.norm(psi0)
    alpha = np.zeros(m)
    beta = np.zeros(max(m - 1, 0))
    v_prev = psi0.copy()
    w = np.empty_like(psi0)
    v_curr = np.empty_like(psi0)
    for j in range(1, m + 1):
        applyH_(w, v_prev, model)
        alpha[j - 1] = np.vdot(v_prev, w).real
        if j == 1:
            w -= alpha[j - 1] * v_prev
        else:
            w -= alpha[j - 1] * v_prev + beta[j - 2] * v_curr
        if j < m:
            beta[j - 1] = np.linalg.norm(w)
            if beta[j - 1] < tol:
                alpha = alpha[:j]
                beta = beta[:j - 1]
                break
            v_curr, v_prev = v_prev, w / beta[j - 1]
    actual_m = len(alpha)
    if actual_m < m:
        beta = beta[:actual_m - 1]
    evals = _eigvals_symtri(alpha, beta)
    return float(evals.min()), float(evals.max())


def lanczos_groundstate(applyH_, model: Model, lanc_m: int = 100, tol: float = 1e-12,
                        orthogonalize_tol: float = 1e-10,
                        rng: Optional[np.random.Generator] = None,
                        v0: Optional[np.ndarray] = None, return_tridiag: bool = False):
    """Lanczos.jl:87-181 (full reorthogonalisation + the second check pass).
    The sign of the returned Ritz vector is LAPACK's choice; compare up to sign."""
    N = len(model)
    m = min(lanc_m, N)
    if v0 is None:
        rng = rng or np.random.default_rng()
        psi0 = rng.standard_normal(N)
    else:
        psi0 = np.array(v0, dtype=np.float64)
    psi0 = psi0 / np.linalg.norm(psi0)
    alpha = np.zeros(m)
    beta = np.zeros(max(m - 1, 0))
    V = np.empty((N, m), order="F")
    V[:, 0] = psi0
    w = np.empty_like(psi0)
    m_actual = m
    vj_buf = np.empty(N)
    for j in range(1, m + 1):
        vj_buf[:] = V[:, j - 1]
        applyH_(w, vj_buf, model)
        if j > 1:                                              # :116-122
            for k in range(1, j):
                vk = V[:, k - 1]
                coeff = np.dot(vk, w)
                w -= coeff * vk
        alpha[j - 1] = np.dot(V[:, j - 1], w)                  # :124
        if j == 1:
            w[:] = w - alpha[j - 1] * V[:, j - 1]
        else:
            w[:] = w - alpha[j - 1] * V[:, j - 1] - beta[j - 2] * V[:, j - 2]
        if j < m:
            beta[j - 1] = np.linalg.norm(w)
            if beta[j - 1] < tol:
                m_actual = j
                break
            for k in range(1, j + 1):                          # :142-153
                vk = V[:, k - 1]
                overlap = abs(np.dot(vk, w / beta[j - 1]))
                if overlap > orthogonalize_tol:
                    w -= np.dot(vk, w) * vk
                    beta[j - 1] = np.linalg.norm(w)
                    if beta[j - 1] < tol:
                        m_actual = j
                        break
            V[:, j] = w / beta[j - 1]
    a_act = alpha[:m_actual]
    b_act = beta[:min(m_actual - 1, len(beta))]
    evals, evecs = _eigen_symtri(a_act, b_act)
    idx = int(np.argmin(evals))
    Emin = float(evals[idx])
    y = evecs[:, idx]
    psi_gs = V[:, :m_actual] @ y
    psi_gs /= np.linalg.norm(psi_gs)
    if return_tridiag:
        return Emin, psi_gs, a_act.copy(), b_act.copy()
    return Emin, psi_gs


def lanczos_tridiag(applyH_, model: Model, v: np.ndarray, lanc_m: int = 100, tol: float = 1e-12):
    """Lanczos.jl:196-246."""
    v = np.asarray(v, dtype=np.complex128)
    n = len(v)
    m = min(lanc_m, n)
    V = [None] * m
    alpha = np.zeros(m)
    beta = np.zeros(max(m - 1, 0))
    w = np.zeros(n, dtype=np.complex128)
    normv = float(np.linalg.norm(v))
    if normv == 0:
        raise RuntimeError("starting vector has zero norm")
    V[0] = v.copy() / normv
    m_eff = m
    for j in range(1, m):
        applyH_(w, V[j - 1], model)
        alpha[j - 1] = np.vdot(V[j - 1], w).real
        w -= alpha[j - 1] * V[j - 1]
        if j > 1:
            w -= beta[j - 2] * V[j - 2]
        beta[j - 1] = np.linalg.norm(w)
        if beta[j - 1] < tol:
            m_eff = j
            break
        V[j] = (w / beta[j - 1]).copy()
    if m_eff == m:
        applyH_(w, V[m - 1], model)
        alpha[m - 1] = np.vdot(V[m - 1], w).real
    else:
        alpha = alpha[:m_eff]
        beta = beta[:m_eff - 1]
    return alpha, beta, normv


def estimate_energy_bounds(applyH_, model: Model, lanc_m: int = 80,
                           rng: Optional[np.random.Generator] = None):
    """Lanczos.jl:255-271 (the reference uses the global RNG; `rng` is an
    extension so tests can be reproducible)."""
    _, Emax = lanczos_extremal(applyH_, model, lanc_m=lanc_m, rng=rng)

    def apply_H_neg_(out, psi, mdl):
        applyH_(out, psi, mdl)
        np.negative(out, out=out)
        return out

    _, Emax_neg = lanczos_extremal(apply_H_neg_, model, lanc_m=lanc_m, rng=rng)
    return -Emax_neg, Emax


# ------------------------------------------------------------ Observables.jl

def _sz_table(model: Model) -> np.ndarray:
    """sz_value(bit_at(state, i)) for every state and site: [N, L]."""
    st = np.array(model.states, dtype=np.uint64)
    return np.where((st[:, None] >> np.arange(model.L, dtype=np.uint64)[None, :]) & np.uint64(1) == 1, 0.5, -0.5)


def magnetization_per_site(psi: np.ndarray, model: Model) -> np.ndarray:
    """Observables.jl:14-37: mags[i] = sum_idx abs2(psi[idx]) * sz_value(bit_at(state, i))."""
    return (np.abs(psi) ** 2) @ _sz_table(model)


def connected_correlations(psi: np.ndarray, model: Model) -> np.ndarray:
    """Observables.jl:43-95: the full SzSz matrix and S_i, then C_r = (1/L) sum_i SzSz[i, mod1(i+r, L)] - S_i[i] S_i[j]."""
    L = model.L
    sz = _sz_table(model)
    amp2 = np.abs(psi) ** 2
    S_i = amp2 @ sz                                        # :64
    SzSz = sz.T @ (amp2[:, None] * sz)                     # :65-67
    C_r = np.zeros(L)
    for r in range(L):                                     # :84-92
        tmp = 0.0
        for i in range(L):
            j = (i + r) % L                                # mod1(i+r, L) with 1-based i
            tmp += SzSz[i, j] - S_i[i] * S_i[j]
        C_r[r] = tmp / L
    return C_r


def structure_factor_Sq(psi: np.ndarray, model: Model) -> dict:
    """Observables.jl:101-109."""
    S_q = np.fft.fft(connected_correlations(psi, model))
    return {2 * np.pi * n / model.L: float(S_q[n].real) for n in range(model.L)}


# ------------------------------------------------------------- LanczosSqw.jl

def spectral_from_tridiagonal(alpha, beta, norm_phi, E0, w_range, eta=0.05, broaden="lorentz"):
    """LanczosSqw.jl:18-43."""
    theta, Q = _eigen_symtri(alpha, beta)
    wts = np.abs(Q[0, :]) ** 2 * norm_phi ** 2
    w_range = np.asarray(w_range, dtype=np.float64)
    shifted = w_range[:, None] - (theta - E0)[None, :]
    if broaden == "lorentz":
        Lmat = (1 / np.pi) * (eta / (shifted ** 2 + eta ** 2))
        return Lmat @ wts
    elif broaden == "gauss":
        pref = 1 / (np.sqrt(2 * np.pi) * eta)
        return (pref * np.exp(-(shifted ** 2) / (2 * eta ** 2))) @ wts
    raise RuntimeError(f"unknown broadening: {broaden}")


def lanczos_sqw(psi0, model: Model, q_list, w_range, lanc_m=200, eta=0.05, broaden="lorentz"):
    """LanczosSqw.jl:49-80.  E0 = real(dot(conj(psi0c), H psi0c)) = sum psi_i (H psi)_i."""
    psi0c = np.asarray(psi0).astype(np.complex128)
    tmp = np.zeros_like(psi0c)
    apply_H_(tmp, psi0c, model)
    E0 = float(np.sum(psi0c * tmp).real)
    S = np.zeros((len(q_list), len(w_range)))
    for iq, q in enumerate(q_list):
        phi = Sz_q_vector(model, psi0c, float(q))
        if np.linalg.norm(phi) == 0:
            continue
        a, b, nphi = lanczos_tridiag(apply_H_, model, phi, lanc_m=lanc_m)
        S[iq, :] = spectral_from_tridiagonal(a, b, nphi, E0, w_range, eta=eta, broaden=broaden)
    return S


# ---------------------------------------------------------------- KPM_Sqw.jl

def _rescaling_from_bounds(E_min, E_max):
    """KPM_Sqw.jl:13-17."""
    return float((E_max - E_min) / (2 * 0.99)), float((E_max + E_min) / 2)


def get_rescaling_params(applyH_, model, lanc_m=80, rng=None):
    """KPM_Sqw.jl:25-28."""
    E_min, E_max = estimate_energy_bounds(applyH_, model, lanc_m=lanc_m, rng=rng)
    return _rescaling_from_bounds(E_min, E_max)


def compute_chebyshev_moments(applyH_, phi, M, a, b, model):
    """KPM_Sqw.jl:95-128."""
    mu = np.zeros(M)
    v_prev = phi.copy()
    v_curr = np.empty_like(phi)
    v_next = np.empty_like(phi)
    mu[0] = np.vdot(phi, v_prev).real
    apply_rescaled_H_(v_curr, v_prev, applyH_, model, a, b)
    mu[1] = np.vdot(phi, v_curr).real
    for m in range(2, M):
        apply_rescaled_H_(v_next, v_curr, applyH_, model, a, b)
        v_next[:] = 2.0 * v_next - v_prev
        mu[m] = np.vdot(phi, v_next).real
        nv = np.linalg.norm(v_next)
        if nv > 1e3:
            v_next /= nv
        v_prev, v_curr, v_next = v_curr, v_next, v_prev
    return mu


def get_kernel(M, kernel):
    """KPM_Sqw.jl:131-145."""
    g = np.ones(M)
    if kernel == "jackson":
        for n in range(M):
            g[n] = ((M - n + 1) * np.cos(np.pi * n / (M + 1))
                    + np.sin(np.pi * n / (M + 1)) / np.tan(np.pi / (M + 1))) / (M + 1)
    elif kernel == "lorentz":
        lam = 3.0
        for n in range(M):
            g[n] = np.sinh(lam * (1 - n / M)) / np.sinh(lam)
    return g


def kpm_sw(phi, applyH_, model, w_range, a, b, E0, kpm_m=200, kernel="jackson"):
    """KPM_Sqw.jl:34-93."""
    mu = compute_chebyshev_moments(applyH_, phi, kpm_m, a, b, model)
    mu = mu * get_kernel(kpm_m, kernel)
    S = np.zeros(len(w_range))
    for iw, w in enumerate(w_range):
        x = (w + E0 - b) / a
        if abs(x) >= 1.0:
            S[iw] = 0.0
            continue
        T = np.zeros(kpm_m)
        T[0] = 1.0
        if kpm_m >= 2:
            T[1] = x
        for n in range(2, kpm_m):
            T[n] = 2.0 * x * T[n - 1] - T[n - 2]
        sum_val = mu[0] * T[0]
        for n in range(1, kpm_m):
            sum_val += 2.0 * mu[n] * T[n]
        denom = np.pi * np.sqrt(1.0 - x ** 2)
        S[iw] = max(0.0, sum_val / (a * denom))
    return S


def kpm_sqw(psi0, model, q_list, w_range, a=None, b=None, kpm_m=200, kernel="jackson", rng=None):
    """KPM_Sqw.jl:191-256."""
    psi0c = np.asarray(psi0).astype(np.complex128)
    S = np.zeros((len(q_list), len(w_range)))
    tmp = np.empty_like(psi0c)
    apply_H_(tmp, psi0c, model)
    E0 = float(np.vdot(psi0c, tmp).real)
    if a is None or b is None:
        a, b = get_rescaling_params(apply_H_, model, rng=rng)
    for iq, q in enumerate(q_list):
        phi = Sz_q_vector(model, psi0c, float(q))
        norm_phi = np.linalg.norm(phi)
        if norm_phi == 0:
            continue
        phi = phi / norm_phi
        Sq = kpm_sw(phi, apply_H_, model, w_range, a=a, b=b, E0=E0, kpm_m=kpm_m, kernel=kernel)
        S[iq, :] = norm_phi ** 2 * Sq
    return S


# ---------------------------------------------------- TimeEvolution/Krylov.jl

def krylov_time_evolve(psi0, dt, applyH_, model, kry_m=30):
    """Krylov.jl:136-192.  Complex dot for alpha; general `eigen(Matrix(TR))`:
    Julia dispatches to the Hermitian solver only when the matrix is exactly
    Hermitian, otherwise to geev; numpy's eigh/eig are used the same way."""
    psi0 = np.asarray(psi0)
    T = psi0.dtype
    n = len(psi0)
    V = [None] * kry_m
    alpha = np.zeros(kry_m, dtype=np.complex128)
    beta = np.zeros(max(kry_m - 1, 0), dtype=np.complex128)
    w = np.zeros(n, dtype=T)
    norm0 = np.linalg.norm(psi0)
    if norm0 == 0:
        return psi0.copy()
    V[0] = psi0.copy() / norm0
    m_eff = kry_m
    for j in range(1, kry_m + 1):
        applyH_(w, V[j - 1], model)
        alpha[j - 1] = np.vdot(V[j - 1], w)
        w -= (alpha[j - 1] * V[j - 1]).astype(T) if T == np.float64 else alpha[j - 1] * V[j - 1]
        if j > 1:
            w -= (beta[j - 2].real * V[j - 2]) if T == np.float64 else beta[j - 2] * V[j - 2]
        if j < kry_m:
            beta[j - 1] = np.linalg.norm(w)
            if abs(beta[j - 1]) < 1e-14:
                m_eff = j
                alpha = alpha[:m_eff]
                beta = beta[:m_eff - 1]
                V = V[:m_eff]
                break
            V[j] = (w / beta[j - 1].real).copy()
    TR = (np.diag(alpha[:m_eff]) + np.diag(beta[:m_eff - 1], 1) + np.diag(beta[:m_eff - 1], -1))
    if np.array_equal(TR, TR.conj().T):
        D, Q = np.linalg.eigh(TR)
    else:
        D, Q = np.linalg.eig(TR)
    U_T = Q @ np.diag(np.exp(-1j * D * dt)) @ Q.conj().T
    e1 = np.zeros(m_eff, dtype=np.complex128)
    e1[0] = norm0
    y = U_T @ e1
    psit = np.zeros(n, dtype=np.complex128)
    for k in range(m_eff):
        psit += y[k] * V[k]
    psit /= np.linalg.norm(psit)
    return psit


# ------------------------------------------------- TimeEvolution/Chebyshev.jl

_MINUS_I_POW = (1.0 + 0j, -1j, -1.0 + 0j, 1j)       # (-im)^k is exact integer arithmetic


def chebyshev_coefficients(dt, cheb_n, Ebounds):
    """Chebyshev.jl:70-79."""
    from scipy.special import jv
    E_min, E_max = Ebounds
    a = (E_max - E_min) / (2 * 0.9999)
    b = (E_max + E_min) / 2
    phase_factor = np.exp(-1j * b * dt)
    c = np.empty(cheb_n, dtype=np.complex128)
    for k in range(cheb_n):
        delta_k0 = 1.0 if k == 0 else 0.0
        c[k] = (2 - delta_k0) * _MINUS_I_POW[k % 4] * jv(k, a * dt) * phase_factor
    return c, a, b


def chebyshev_time_evolve(psi0, dt, applyH_, model, cheb_n=100, Ebounds=(-1.0, 1.0)):
    """Chebyshev.jl:61-124.  psi0 must be complex (a real psi0 raises
    InexactError in the reference at `ws.psi_t[i] += c[1]*...`)."""
    assert cheb_n >= 1, "cheb_n must be >= 1"
    psi0 = np.asarray(psi0)
    if psi0.dtype != np.complex128:
        raise TypeError("InexactError: real psi0 cannot hold complex Chebyshev sums")
    c, a, b = chebyshev_coefficients(dt, cheb_n, Ebounds)
    phi_prev = psi0.copy()
    phi_curr = np.empty_like(psi0)
    phi_next = np.empty_like(psi0)
    apply_rescaled_H_(phi_curr, phi_prev, applyH_, model, a, b)
    psi_t = np.zeros_like(psi0)
    psi_t += c[0] * phi_prev
    if cheb_n >= 2:
        psi_t += c[1] * phi_curr
    if cheb_n == 1:
        return psi_t.copy()
    for k in range(2, cheb_n):
        apply_rescaled_H_(phi_next, phi_curr, applyH_, model, a, b)
        phi_next[:] = 2 * phi_next - phi_prev
        psi_t += c[k] * phi_next
        phi_prev, phi_curr, phi_next = phi_curr, phi_next, phi_prev
    return psi_t.copy()


# ------------------------------------------- TimeEvolution/KPM.jl (site-resolved KPM)

def site_sz_operator(i: int):
    """create_spin_operator(i, :z) as TimeEvolution/KPM.jl:205-206 uses it: psi -> S^z_i psi (1-based site i).  The
    reference builds it in Operators.jl (out of the hot-path scope); S^z is diagonal: s_i(state) * psi."""
    def op(psi, model):
        st = np.array(model.states, dtype=np.uint64)
        s = np.where((st >> np.uint64(i - 1)) & np.uint64(1) == np.uint64(1), 0.5, -0.5)
        return (s * np.asarray(psi)).astype(np.complex128)
    return op


def kpm_get_rescaling_params(applyH_, model, lanc_m=80, rng=None):
    """TimeEvolution/KPM.jl:45-49: a = (E_max - E_min) / 2 * 0.9, b = (E_max + E_min) / 2."""
    E_min, E_max = estimate_energy_bounds(applyH_, model, lanc_m=lanc_m, rng=rng)
    return (E_max - E_min) / 2 * 0.9, (E_max + E_min) / 2


def get_jackson_kernel(n: int):
    """TimeEvolution/KPM.jl:170-177."""
    g = np.zeros(n)
    for k in range(n):
        d = np.pi / (n + 1)
        g[k] = ((n - k + 1) * np.cos(d * k) + np.sin(d * k) / np.tan(d)) / (n + 1)
    return g


def evaluate_chebyshev_series(mu, x, a):
    """TimeEvolution/KPM.jl:184-206."""
    n = len(mu)
    if abs(x) >= 1.0:
        return 0.0
    total = mu[0] * 1.0
    if n > 1:
        total += mu[1] * x
    T_prev, T_curr = 1.0, x
    for k in range(2, n):
        T_next = 2 * x * T_curr - T_prev
        total += mu[k] * T_next
        T_prev, T_curr = T_curr, T_next
    return total / (np.pi * np.sqrt(1 - x ** 2)) * (2 / a)


def compute_cross_chebyshev_moments(chi, phi, n, a, b, applyH_, model):
    """TimeEvolution/KPM.jl:121-165.  `dot(conj(chi), x)` = sum_i chi_i x_i (Julia's dot conjugates its first argument)."""
    moments = np.zeros(n)
    phi = np.asarray(phi)
    chi = np.asarray(chi)
    phi_prev = phi.copy()
    norm_phi = np.linalg.norm(phi)
    phi_prev = phi_prev / norm_phi
    phi_curr = np.empty_like(phi_prev)
    temp = np.empty_like(phi_prev)
    apply_rescaled_H_(phi_curr, phi_prev, applyH_, model, a, b)

    def to_real(z):                                                   # assignment of a Complex to a Float64 slot
        z = complex(z)
        if z.imag != 0.0:
            raise TypeError("InexactError")
        return z.real

    moments[0] = to_real(np.sum(chi * phi_prev) * norm_phi)
    if n > 1:
        moments[1] = to_real(np.sum(chi * phi_curr) * norm_phi)
    for k in range(2, n):
        apply_rescaled_H_(temp, phi_curr, applyH_, model, a, b)
        phi_next = 2 * temp - phi_prev
        moments[k] = np.real(np.sum(chi * phi_next)) * norm_phi
        phi_prev, phi_curr = phi_curr, phi_next
    return moments


def kpm_dynamical_correlation(psi, operator_A, operator_B, w_range, applyH_, model, n=300, eps=0.1, a=None, b=None):
    """TimeEvolution/KPM.jl:74-118."""
    if a is None or b is None:
        a, b = kpm_get_rescaling_params(applyH_, model, lanc_m=n)
    phi = operator_B(psi, model)
    chi = operator_A(psi, model)
    mu = compute_cross_chebyshev_moments(chi, phi, n, a, b, applyH_, model)
    mu = mu * get_jackson_kernel(n)
    S = np.zeros(len(w_range))
    for i, w in enumerate(w_range):
        S[i] = evaluate_chebyshev_series(mu, (w - b) / a, a)
    return np.maximum(S, 0.0)


# -------------------------------------------------------------- PublicAPI.jl

def groundstate(model, method="lanczos", **kw):
    """PublicAPI.jl:25-35."""
    if method == "lanczos":
        return lanczos_groundstate(apply_H_, model, **kw)
    raise ValueError(f"unsupported ground-state method: {method}")


def time_evolve(model, psi0, t, method="krylov", Ebounds=None, **kw):
    """PublicAPI.jl:50-88."""
    if method == "krylov":
        return krylov_time_evolve(psi0, float(t), apply_H_, model, **kw)
    elif method == "chebyshev":
        bounds = estimate_energy_bounds(apply_H_, model) if Ebounds is None else Ebounds
        return chebyshev_time_evolve(psi0, float(t), apply_H_, model, Ebounds=bounds, **kw)
    raise ValueError(f"unsupported time-evolution method: {method}")


def structure_factor(model, psi):
    """PublicAPI.jl:94-106."""
    return structure_factor_Sq(psi, model)


def dynamical_structure_factor(model, psi0, q, w, method="lanczos", **kw):
    """PublicAPI.jl:122-155."""
    q_list = np.asarray(q, dtype=np.float64)
    w_range = np.asarray(w, dtype=np.float64)
    if method == "lanczos":
        return lanczos_sqw(psi0, model, q_list, w_range, **kw)
    elif method == "kpm":
        return kpm_sqw(psi0, model, q_list, w_range, **kw)
    raise ValueError(f"unsupported dynamical structure-factor method: {method}")


# ---------------------------------------------------- bench-input generator

def fill_seeded(N: int, seed: int, cplx: bool = False, first: int = 0) -> np.ndarray:
    """SURVEY.md 8(d) counter-based synthetic psi (unnormalised)."""
    v = np.empty(N, dtype=np.complex128 if cplx else np.float64)
    lib().orc_fill_seeded(_ptr(v), first, N, seed, 1 if cplx else 0)
    return v


def row_seeded_f64(model_args, state: int, seed: int, scale: float = 1.0) -> float:
    """(H psi)[rank(state)] for the seeded psi without materialising anything."""
    L, nup, hop, zz, fld = model_args
    ha, za = _bond_array(hop), _bond_array(zz)
    fld = np.ascontiguousarray(fld, dtype=np.float64)
    return lib().orc_row_seeded_f64(L, -1 if nup is None else nup,
                                    ctypes.cast(ha, ctypes.c_void_p), len(hop),
                                    ctypes.cast(za, ctypes.c_void_p), len(zz), _ptr(fld),
                                    state, seed, scale)
