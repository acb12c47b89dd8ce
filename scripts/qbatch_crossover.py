"""Where does the q-batched multi-vector recurrence stop paying?  lanczos_sqw over the momenta of the chain, batched against
the per-momentum loop, for a few chain lengths (single GPU).  Sets Q_BATCH_AUTO_MAX_DIM in spindyn/api.py."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spindynamics.jl_b200"))
import numpy as np
import spindyn as sd

w = np.linspace(0.0, 4.0, 50)
for L in [int(x) for x in sys.argv[1:]] or [16, 20, 22, 24]:
    m = sd.XXZChain(L, nup=L // 2)
    psi = m.vector(np.float64).fill_seeded(3, 1.0 / np.sqrt(m.dim / 3.0))
    q = sd.momenta(m)
    res = {}
    for batch in (True, False):
        sd.lanczos_sqw(psi, m, q, w, lanc_m=4, q_batch=batch)
        m.ctx.sync()
        t0 = time.perf_counter()
        S = sd.lanczos_sqw(psi, m, q, w, lanc_m=40, q_batch=batch)
        m.ctx.sync()
        res[batch] = (time.perf_counter() - t0, S)
    d = float(np.linalg.norm(res[True][1] - res[False][1]) / np.linalg.norm(res[False][1]))
    print(f"L={L} N={m.dim} nq={len(q)}: batched {res[True][0] * 1e3:.1f} ms, loop {res[False][0] * 1e3:.1f} ms, ratio loop/batched "
          f"{res[False][0] / res[True][0]:.2f}, rel diff {d:.1e}", flush=True)
    del psi, m
