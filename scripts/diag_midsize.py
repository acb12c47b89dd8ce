"""Where does a Lanczos step go at mid sizes?  Times the pieces of the per-momentum path of lanczos_sqw at a few L."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import numpy as np
import spindyn as sd

lib, check = sd.lib(), sd._lib.check


def timed(ctx, fn, reps):
    fn()
    ctx.sync()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    ms = ctx.timer_stop()
    wall = (time.perf_counter() - t0) * 1e3
    return ms / reps, wall / reps, (ctx.launch_count() - l0) / reps


for L in [int(x) for x in sys.argv[1:]] or [22, 24, 26]:
    m = sd.XXZChain(L, nup=L // 2)
    ctx = m.ctx
    for dt in (np.float64, np.complex128):
        x = m.vector(dt).fill_seeded(3, 1.0 / np.sqrt(m.dim / 3.0))
        y = m.vector(dt)
        r = sd._lib.SdComplex()
        a = timed(ctx, lambda: sd.apply_H_(y, x, m), 50)
        b = timed(ctx, lambda: check(lib.sd_apply_H_dot(m._h, y._h, x._h, ctypes.byref(r))), 50)
        print(f"L={L} {np.dtype(dt).name}: apply {a[0]:.3f} ms (wall {a[1]:.3f}, {a[2]:.0f} launches); apply+dot with fetch {b[0]:.3f} ms (wall {b[1]:.3f}, {b[2]:.0f} launches)", flush=True)
        if dt == np.complex128:
            phi = m.vector(dt)
            n2 = ctypes.c_double()
            c = timed(ctx, lambda: check(lib.sd_szq(m._h, phi._h, x._h, 1.0, ctypes.byref(n2))), 20)
            d = timed(ctx, lambda: sd.lanczos_tridiag(sd.apply_H_, m, phi, lanc_m=40), 5)
            print(f"L={L}: szq {c[0]:.3f} ms (wall {c[1]:.3f}); lanczos_tridiag(40) {d[0]:.3f} ms = {d[0] / 40:.3f} per step (wall {d[1]:.3f}, {d[2]:.0f} launches)", flush=True)
        del x, y
    del m
