"""The copy engine behind sd_vec_upload_async / sd_vec_download_async (sd_api.cu): pinned host buffers, two copy streams,
transfers cut into tile-aligned chunks whose layout permutes overlap the next chunk, uploads running ahead of the kernels
and overlapping a preceding download.  Checked for what a caller can observe: after sd_ctx_sync every downloaded buffer
holds exactly what the synchronous path produces, whatever was in flight at the same time."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from conftest import sd  # noqa: E402


def _sync_apply(m, x):
    d = m.to_device(x)
    o = m.vector(x.dtype)
    sd.apply_H_(o, d, m)
    return o.to_host()


@pytest.mark.parametrize("L,nup,dtype", [(26, 13, np.float64), (24, 12, np.complex128), (16, 8, np.float64)])
def test_async_copies_in_flight_equal_the_synchronous_path(L, nup, dtype):
    """L = 26 f64 (83 MB per vector) and L = 24 c128 are cut into 8 chunks, L = 16 is one chunk.  Three applies in flight
    through two (psi, out) pairs and three distinct host inputs; a synchronous upload / download and a kernel that uses
    the staging buffer (szq on the other dtype is not needed: sd_vec_convert does) are interleaved on purpose."""
    m = sd.XXZChain(L, Jxy=0.9, Jz=1.1, hz=0.1, nup=nup)
    assert m.info["kernel_path"] == "block"
    N = m.dim
    rng = np.random.default_rng(L)
    lib, check = sd.lib(), sd._lib.check
    hin = [sd.PinnedBuffer(N, dtype) for _ in range(3)]
    hout = [sd.PinnedBuffer(N, dtype) for _ in range(3)]
    for h in hin:
        h.array[:] = rng.standard_normal(N)
        if dtype == np.complex128:
            h.array[:] += 1j * rng.standard_normal(N)
    for h in hout:
        h.array[:] = np.nan
    want = [_sync_apply(m, h.array.copy()) for h in hin]
    pairs = [(m.vector(dtype), m.vector(dtype)) for _ in range(2)]
    for s in range(3):
        p, o = pairs[s % 2]
        check(lib.sd_vec_upload_async(p._h, hin[s]._p))
        check(lib.sd_apply_H(m._h, o._h, p._h))
        check(lib.sd_vec_download_async(o._h, hout[s]._p))
        if s == 1:                                                  # a synchronous round trip in the middle of the pipeline
            z = m.to_device(hin[0].array.copy())
            assert np.array_equal(z.to_host(), hin[0].array)
            if dtype == np.float64:                                 # and a compute-stream user of the upload staging buffer
                zc = m.vector(np.complex128)
                check(lib.sd_vec_convert(zc._h, z._h))
                assert np.array_equal(zc.to_host().real, hin[0].array)
    m.ctx.sync()
    for s in range(3):
        assert np.array_equal(hout[s].array, want[s]), s
    # a second round re-using everything (events, staging buffers, pairs) and timed with the stopwatch that joins the streams
    for h in hout:
        h.array[:] = np.nan
    m.ctx.timer_start()
    for s in range(3):
        p, o = pairs[(s + 1) % 2]
        check(lib.sd_vec_upload_async(p._h, hin[2 - s]._p))
        check(lib.sd_apply_H(m._h, o._h, p._h))
        check(lib.sd_vec_download_async(o._h, hout[s]._p))
    assert m.ctx.timer_stop() > 0.0
    for s in range(3):
        assert np.array_equal(hout[s].array, want[2 - s]), s
    del pairs
    for h in hin + hout:
        h.free()


def test_apply_H_host_entry_point_uses_the_engine():
    """sd_apply_H_host (what apply_H!(::Vector, ::Vector, ::Model) binds to) with pageable numpy buffers."""
    m = sd.XXZChain(20, nup=10)
    x = np.random.default_rng(3).standard_normal(m.dim)
    out = np.empty_like(x)
    sd.apply_H_(out, x, m)
    assert np.array_equal(out, _sync_apply(m, x))
