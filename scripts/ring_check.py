"""GPU check of the ring kernel (sd_blkr.h, SD_BLK_RING=1, f64) against the oracle and against the standard block
kernel, then timing.  Run it under `timeout`: the kernel has a watchdog (a protocol error traps instead of hanging),
but it has never run on a GPU before round 2.
    SD_BLK_RING=1 timeout 300 python scripts/ring_check.py 28 32
"""
import os
import sys

os.environ.setdefault("SD_BLK_RING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import numpy as np  # noqa: E402
import spindyn as sd  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


ok = True
print("SD_BLK_RING =", os.environ["SD_BLK_RING"], flush=True)
for (L, nup) in [(16, 8), (16, 2), (16, 14), (16, 0), (16, 16), (16, 1), (16, 15), (17, 8), (18, 9), (19, 9), (20, 10), (20, 3), (21, 11), (22, 11)]:
    m = sd.XXZChain(L, Jxy=0.7, Jz=1.3, hz=0.2, nup=nup)
    om = orc.XXZChain(L, Jxy=0.7, Jz=1.3, hz=0.2, nup=nup)
    assert m.info["kernel_path"] == "block", m.info
    rng = np.random.default_rng(L * 100 + nup)
    psi = rng.standard_normal(m.dim)
    ref = np.empty_like(psi)
    orc.apply_H_(ref, psi, om)
    out = np.empty_like(psi)
    sd.apply_H_(out, psi, m)
    e = rel(out, ref)
    # fused <psi, H psi> (Lanczos alpha) and a Lanczos ground state through the fused epilogues
    d, o = m.to_device(psi), m.vector(np.float64)
    import ctypes
    r = sd._lib.SdComplex()
    sd._lib.check(sd.lib().sd_apply_H_dot(m._h, o._h, d._h, ctypes.byref(r)))
    de = abs(r.re - psi @ ref) / max(1.0, abs(psi @ ref))
    e2 = rel(o.to_host(), ref)
    good = e < 1e-13 and e2 < 1e-13 and de < 1e-12
    ok &= good
    print(f"L={L} nup={nup}: apply {e:.2e}  apply+dot {e2:.2e} dot {de:.2e}  {'OK' if good else 'FAIL'}", flush=True)
for L in (16, 18):
    m = sd.XXZChain(L, nup=L // 2)
    om = orc.XXZChain(L, nup=L // 2)
    v0 = np.random.default_rng(3).standard_normal(m.dim)
    E0, _ = sd.groundstate(m, lanc_m=60, v0=v0)
    E0r, _ = orc.groundstate(om, lanc_m=60, v0=v0)
    print(f"L={L} groundstate E0={E0:.12f} (oracle {E0r:.12f})", flush=True)
    ok &= abs(E0 - E0r) < 1e-9
print("ALL OK" if ok else "SOME FAILED", flush=True)
for L in [int(x) for x in sys.argv[1:]]:
    m = sd.XXZChain(L, nup=L // 2)
    x = m.vector(np.float64).fill_seeded(1)
    y = m.vector(np.float64)
    for _ in range(3):
        sd.apply_H_(y, x, m)
    m.ctx.sync()
    m.ctx.timer_start()
    for _ in range(10):
        sd.apply_H_(y, x, m)
    ms = m.ctx.timer_stop() / 10
    # Hermiticity at full size: <x, H y'> = <H x, y'> with y' = H x
    print(f"L={L} ring={os.environ['SD_BLK_RING']} f64: {ms:.3f} ms/apply  frac={16 * m.dim / ms / 1e6 / 6552.3:.3f}  "
          f"checksum={y.dot(x).real:.12e}", flush=True)
sys.exit(0 if ok else 1)
