"""Debug: per-phase cycle sums of the tiled kernel (needs libspindyn_cuda_timing.so, built with -DSD_PHASE_TIMING)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import spindyn._lib as L
L.LIB_PATH = os.path.join(ROOT, "spindynamics.jl_b200", "libspindyn_cuda_timing.so")
import spindyn as sd
Lc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = sd.XXZChain(Lc, nup=Lc // 2)
psi = m.vector().fill_seeded(1, 1e-4); out = m.vector()
lib = sd.lib()
lib.sd_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros(16, dtype=np.uint64)
sd.apply_H_(out, psi, m); m.ctx.sync()
lib.sd_debug_phase_cycles(m.ctx._h, buf.ctypes.data_as(ctypes.c_void_p), 1)
m.ctx.timer_start()
for _ in range(3): sd.apply_H_(out, psi, m)
ms = m.ctx.timer_stop() / 3
lib.sd_debug_phase_cycles(m.ctx._h, buf.ctypes.data_as(ctypes.c_void_p), 1)
names = ["phase0(thread0)", "bar0 wait", "phase1(thread0)", "bar1 wait", "phase2(thread0)", "bar2 wait", "phase3(thread0)", "-", "p1: perm+cp.async", "p1: far streams", "p1: near streams", "p1: crossing", "p1: cp.async wait", "-", "-", "-"]
tot = buf[:8].sum()
ntiles = 1 << (Lc - m.info["tile_sites"])
print(f"L={Lc} ms/apply={ms:.3f} tiles={ntiles} threads={os.environ.get('SD_TILE_THREADS','512')} B={m.info['tile_sites']}")
for n, c in zip(names, buf):
    print(f"  {n:18s} {100.0 * c / tot:5.1f}%  {c / 3 / ntiles:9.0f} cycles/tile")
print(f"  total {tot / 3 / ntiles:.0f} cycles/tile = {tot / 3 / ntiles / 1.965e3:.2f} us")
