"""Handles over libspindyn_cuda: Context, Model (mirror of SpinModel.Model),
DeviceVector and VecSet.  Host-side twin of the Julia wrapper types in
INTEGRATION.md.  Everything that touches vector data runs on the GPU through
the C ABI; numpy is only used for host buffers and the small dense problems the
reference also solves on the host (tridiagonal eigenproblems etc.).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import SD_C128, SD_F64, SdBond, SdComplex, check, lib

_DEFAULT_CTX = None


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def device_count() -> int:
    n = ctypes.c_int(0)
    rc = lib().sd_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


class Context:
    """One GPU + one stream; rank `rank` of `world` cooperating processes when
    created through `Context.for_rank` (one process per GPU, NCCL + CUDA IPC)."""

    def __init__(self, device: int = 0, _handle=None):
        if _handle is None:
            h = ctypes.c_void_p()
            check(lib().sd_ctx_create(int(device), ctypes.byref(h)))
            _handle = h
        self._h = _handle
        self.device = int(device)

    @classmethod
    def for_rank(cls, device: int, rank: int, world: int, nccl_id: bytes) -> "Context":
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(bytes(nccl_id), 128)
        check(lib().sd_ctx_create_rank(int(device), int(rank), int(world), buf, ctypes.byref(h)))
        return cls(device, _handle=h)

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        check(lib().sd_nccl_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: Optional[int] = None) -> "Context":
        """One rank per GPU under torchrun: rank 0 makes the NCCL id, torch.distributed
        (any backend) broadcasts it."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", rank))
        if world == 1:
            return cls(device)
        obj = [cls.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        return cls.for_rank(device, rank, world, obj[0])

    @property
    def rank(self) -> int:
        r, w = ctypes.c_int(), ctypes.c_int()
        check(lib().sd_ctx_rank(self._h, ctypes.byref(r), ctypes.byref(w)))
        return r.value

    @property
    def world(self) -> int:
        r, w = ctypes.c_int(), ctypes.c_int()
        check(lib().sd_ctx_rank(self._h, ctypes.byref(r), ctypes.byref(w)))
        return w.value

    def sync(self) -> None:
        check(lib().sd_ctx_sync(self._h))

    def collect(self) -> None:
        """Collective (world > 1): release the shards of vectors that every rank has freed (sd_ctx_collect)."""
        check(lib().sd_ctx_collect(self._h))

    def timer_start(self) -> None:
        check(lib().sd_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = ctypes.c_float()
        check(lib().sd_timer_stop(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def launch_count(self) -> int:
        n = ctypes.c_uint64()
        check(lib().sd_launch_count(self._h, ctypes.byref(n)))
        return int(n.value)

    def close(self) -> None:
        if self._h:
            lib().sd_ctx_free(self._h)
            self._h = None


def default_context() -> Context:
    global _DEFAULT_CTX
    if _DEFAULT_CTX is None:
        _DEFAULT_CTX = Context(0)
    return _DEFAULT_CTX


def set_default_context(ctx: Optional[Context]) -> None:
    global _DEFAULT_CTX
    _DEFAULT_CTX = ctx


def _bond_array(lst):
    arr = (SdBond * max(1, len(lst)))()
    for k, (i, j, J) in enumerate(lst):
        arr[k].i, arr[k].j, arr[k].J = int(i), int(j), float(J)
    return arr


class Model:
    """SpinModel.jl:6-15.  Same fields (L, nup, mode, hopping_list, onsite_field,
    zz_list); `states` / `idxmap` are NOT materialised on the GPU path -- `states`
    is produced on demand by unranking on the device, `idxmap` lookups go through
    `rank_of` (sd_rank)."""

    def __init__(self, L, nup, hopping, onsite_field, zz, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.L = int(L)
        self.nup = None if nup is None else int(nup)
        self.mode = "full" if nup is None else "sector"
        self.hopping_list = [(int(i), int(j), float(J)) for (i, j, J) in hopping]
        self.zz_list = [(int(i), int(j), float(J)) for (i, j, J) in zz]
        self.onsite_field = np.ascontiguousarray(onsite_field, dtype=np.float64)
        if self.onsite_field.shape != (self.L,) and self.L >= 1:
            raise ValueError("onsite_field must have L entries")
        ha, za = _bond_array(self.hopping_list), _bond_array(self.zz_list)
        h = ctypes.c_void_p()
        check(lib().sd_model_create(self.ctx._h, self.L, -1 if nup is None else int(nup),
                                    ctypes.cast(ha, ctypes.c_void_p), len(self.hopping_list),
                                    ctypes.cast(za, ctypes.c_void_p), len(self.zz_list),
                                    _ptr(self.onsite_field), ctypes.byref(h)))
        self._h = h
        n = ctypes.c_uint64()
        check(lib().sd_model_dim(self._h, ctypes.byref(n)))
        self.dim = int(n.value)

    def __len__(self) -> int:
        return self.dim

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.ctx._h:
                lib().sd_model_free(self._h)
                self._h = None
        except Exception:
            pass

    # -- basis (Basis.jl) ---------------------------------------------------
    def unrank(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """states[first : first+count] of build_sector_basis / build_full_basis."""
        if count is None:
            count = self.dim - first
        out = np.empty(count, dtype=np.uint64)
        check(lib().sd_unrank(self._h, int(first), int(count), _ptr(out)))
        return out

    @property
    def states(self) -> np.ndarray:
        return self.unrank(0, self.dim)

    def rank_of(self, states) -> np.ndarray:
        """get(idxmap, s, 0): 1-based index, 0 if absent."""
        s = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1)
        out = np.empty(s.shape[0], dtype=np.int64)
        check(lib().sd_rank(self._h, _ptr(s), s.shape[0], _ptr(out)))
        return out

    # -- sharding -----------------------------------------------------------
    @property
    def local_range(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        check(lib().sd_model_local_range(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def shard_bounds(self, world: int) -> np.ndarray:
        out = np.zeros(world + 1, dtype=np.uint64)
        check(lib().sd_model_shard_bounds(self._h, int(world), _ptr(out)))
        return out

    @property
    def info(self):
        p, t, r = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib().sd_model_info(self._h, ctypes.byref(p), ctypes.byref(t), ctypes.byref(r)))
        return {"kernel_path": {_lib.SD_PATH_TILED: "tiled", _lib.SD_PATH_BLOCK: "block"}.get(p.value, "generic"),
                "tile_sites": t.value, "rank_bits": r.value}

    def set_path(self, path: str) -> None:
        """"block" stores vectors in block layout, "tiled"/"generic" in rank order: switching between
        the two groups is only possible while no DeviceVector of the model is alive."""
        check(lib().sd_model_set_path(self._h, {"generic": 0, "tiled": 1, "block": 2}[path]))

    # -- vectors ------------------------------------------------------------
    def vector(self, dtype=np.float64) -> "DeviceVector":
        return DeviceVector(self, dtype)

    def to_device(self, host: np.ndarray) -> "DeviceVector":
        """Upload the LOCAL shard (whole vector in a single-rank context)."""
        host = np.ascontiguousarray(host)
        if host.dtype not in (np.float64, np.complex128):
            host = host.astype(np.complex128 if np.iscomplexobj(host) else np.float64)
        v = DeviceVector(self, host.dtype)
        if host.shape != (v.local_len,):
            raise ValueError(f"DimensionMismatch: expected {v.local_len} elements, got {host.shape}")
        check(lib().sd_vec_upload(v._h, _ptr(host)))
        return v


def _sd_dtype(dtype) -> int:
    dt = np.dtype(dtype)
    if dt == np.float64:
        return SD_F64
    if dt == np.complex128:
        return SD_C128
    raise TypeError(f"unsupported element type {dt}; use float64 or complex128")


class DeviceVector:
    """A psi buffer resident in HBM (opaque sd_vec handle)."""

    def __init__(self, model: Model, dtype=np.float64, _handle=None, _owned=True):
        self.model = model
        self.dtype = np.dtype(dtype)
        self._owned = _owned
        if _handle is None:
            h = ctypes.c_void_p()
            check(lib().sd_vec_alloc(model._h, _sd_dtype(dtype), ctypes.byref(h)))
            _handle = h
        self._h = _handle
        n = ctypes.c_uint64()
        check(lib().sd_vec_local_len(self._h, ctypes.byref(n)))
        self.local_len = int(n.value)

    def __len__(self):
        return self.model.dim

    def free(self):
        if self._h and self._owned:
            lib().sd_vec_free(self._h)
        self._h = None

    def __del__(self):
        try:
            if self.model.ctx._h:
                self.free()
        except Exception:
            pass

    def upload(self, host: np.ndarray) -> "DeviceVector":
        host = np.ascontiguousarray(host, dtype=self.dtype)
        if host.shape != (self.local_len,):
            raise ValueError(f"DimensionMismatch: expected {self.local_len} elements, got {host.shape}")
        check(lib().sd_vec_upload(self._h, _ptr(host)))
        return self

    def to_host(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.local_len, dtype=self.dtype)
        check(lib().sd_vec_download(self._h, _ptr(out)))
        return out

    def zero(self):
        check(lib().sd_vec_zero(self._h))
        return self

    def set_onehot(self, idx0: int):
        check(lib().sd_vec_set_onehot(self._h, int(idx0)))
        return self

    def get(self, idx0):
        """Elements by 0-based basis rank: (values, present) -- present[i] is False for ranks another shard holds."""
        idx = np.ascontiguousarray(idx0, dtype=np.uint64)
        out = np.zeros(len(idx), dtype=self.dtype)
        present = np.zeros(len(idx), dtype=np.uint8)
        check(lib().sd_vec_get(self._h, _ptr(idx), len(idx), _ptr(out), _ptr(present)))
        return out, present.astype(bool)

    def fill_seeded(self, seed: int, scale: float = 1.0):
        check(lib().sd_vec_fill_seeded(self._h, int(seed), float(scale)))
        return self

    def copy_from(self, src: "DeviceVector"):
        check(lib().sd_vec_copy(self._h, src._h))
        return self

    def convert_from(self, src: "DeviceVector"):
        check(lib().sd_vec_convert(self._h, src._h))
        return self

    def astype(self, dtype) -> "DeviceVector":
        out = DeviceVector(self.model, dtype)
        return out.convert_from(self)

    def scale(self, s):
        s = complex(s)
        check(lib().sd_vec_scale(self._h, SdComplex(s.real, s.imag)))
        return self

    def axpy(self, a, x: "DeviceVector"):
        a = complex(a)
        check(lib().sd_vec_axpy(self._h, SdComplex(a.real, a.imag), x._h))
        return self

    def dot(self, y: "DeviceVector") -> complex:
        """LinearAlgebra.dot(self, y) = sum conj(self_i) y_i."""
        r = SdComplex()
        check(lib().sd_vec_dot(self._h, y._h, ctypes.byref(r)))
        return complex(r.re, r.im)

    def dotu(self, y: "DeviceVector") -> complex:
        r = SdComplex()
        check(lib().sd_vec_dotu(self._h, y._h, ctypes.byref(r)))
        return complex(r.re, r.im)

    def norm(self) -> float:
        r = ctypes.c_double()
        check(lib().sd_vec_norm(self._h, ctypes.byref(r)))
        return float(r.value)


class VecSet:
    """Device-resident Lanczos/Krylov basis (V of Lanczos.jl:104, Krylov.jl:140)."""

    def __init__(self, model: Model, handle, dtype):
        self.model, self._h, self.dtype = model, handle, np.dtype(dtype)

    def __len__(self):
        m = ctypes.c_int()
        check(lib().sd_vecset_size(self._h, ctypes.byref(m)))
        return m.value

    def __getitem__(self, k: int) -> DeviceVector:
        h = ctypes.c_void_p()
        check(lib().sd_vecset_get(self._h, int(k), ctypes.byref(h)))
        return DeviceVector(self.model, self.dtype, _handle=h, _owned=False)

    def lincomb(self, y: Sequence[complex], out: DeviceVector) -> float:
        """out = sum_k y[k] V_k; returns ||out||^2."""
        yc = np.ascontiguousarray(y, dtype=np.complex128)
        n2 = ctypes.c_double()
        check(lib().sd_lincomb(self._h, _ptr(yc), len(yc), out._h, ctypes.byref(n2)))
        return float(n2.value)

    def free(self):
        if self._h:
            lib().sd_vecset_free(self._h)
            self._h = None

    def __del__(self):
        try:
            if self.model.ctx._h:
                self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host buffer for asynchronous copies (bench e2e path)."""

    def __init__(self, n: int, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(n) * self.dtype.itemsize
        p = ctypes.c_void_p()
        check(lib().sd_host_alloc(ctypes.byref(p), self.nbytes))
        self._p = p
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(n))

    def free(self):
        if self._p:
            self.array = None
            lib().sd_host_free(self._p)
            self._p = None
