"""GPU check of the block-layout kernel against the oracle (small sizes) and timing at L=28/32."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import numpy as np
import spindyn as sd
from oracle import oracle as orc

def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
ok = True
for (L, nup) in [(16, 8), (16, 2), (16, 14), (17, 8), (18, 9), (20, 10), (20, 3), (21, 11), (16, 0), (16, 16), (16, 1), (16, 15)]:
    for dtype in (np.float64, np.complex128):
        m = sd.XXZChain(L, Jxy=0.7, Jz=1.3, hz=0.2, nup=nup)
        om = orc.XXZChain(L, Jxy=0.7, Jz=1.3, hz=0.2, nup=nup)
        assert m.info["kernel_path"] == "block", m.info
        rng = np.random.default_rng(L * 100 + nup)
        psi = rng.standard_normal(m.dim).astype(dtype)
        if dtype == np.complex128: psi = psi + 1j * rng.standard_normal(m.dim)
        # round trip of the layout conversion
        d = m.to_device(psi)
        back = d.to_host()
        rt = np.array_equal(back, psi)
        ref = np.empty_like(psi); orc.apply_H_(ref, psi, om)
        out = np.empty_like(psi); sd.apply_H_(out, psi, m)
        e = rel(out, ref)
        good = rt and e < 1e-13
        ok &= good
        print(f"L={L} nup={nup} {np.dtype(dtype).name}: roundtrip={rt} relerr={e:.2e} {'OK' if good else 'FAIL'}", flush=True)
print("ALL OK" if ok else "SOME FAILED")
if len(sys.argv) > 1:
    for L in [int(x) for x in sys.argv[1:]]:
        m = sd.XXZChain(L, nup=L // 2)
        x = m.vector(np.float64).fill_seeded(1)
        y = m.vector(np.float64)
        for _ in range(3): sd.apply_H_(y, x, m)
        m.ctx.sync(); m.ctx.timer_start()
        for _ in range(10): sd.apply_H_(y, x, m)
        ms = m.ctx.timer_stop() / 10
        print(f"L={L} block f64: {ms:.3f} ms/apply  frac={16*m.dim/ms/1e6/6552.3:.3f}", flush=True)
