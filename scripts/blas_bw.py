"""HBM bandwidth of the BLAS-1 kernels through the C ABI on an L = 32 f64 / L = 30 c128 vector (CUDA-event timed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import numpy as np  # noqa: E402
import spindyn as sd  # noqa: E402

ctx = sd.Context(0)
sd.set_default_context(ctx)
for (L, dtype, esz) in [(32, np.float64, 8), (30, np.complex128, 16)]:
    m = sd.XXZChain(L, nup=L // 2, ctx=ctx)
    x = m.vector(dtype).fill_seeded(1, 1e-4)
    y = m.vector(dtype).fill_seeded(2, 1e-4)
    n = m.dim

    def timeit(name, fn, nbytes, reps=10):
        for _ in range(2):
            fn()
        ctx.sync()
        ctx.timer_start()
        for _ in range(reps):
            fn()
        ms = ctx.timer_stop() / reps
        print(f"L={L} {np.dtype(dtype).name} {name}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s ({nbytes / ms / 1e6 / 6552.3:.2f} of the measured HBM peak)", flush=True)

    timeit("axpy  y += a x (2R+1W)", lambda: y.axpy(1e-3, x), 3 * esz * n)
    timeit("dot   <x,y>   (2R)", lambda: x.dot(y), 2 * esz * n)
    timeit("scale x *= s  (1R+1W)", lambda: x.scale(1.0000001), 2 * esz * n)
    timeit("norm  ||x||   (2R same vector)", lambda: x.norm(), 1 * esz * n)
    del x, y, m
