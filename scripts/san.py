"""compute-sanitizer reproducer: block kernel with a non-plain epilogue."""
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/spindynamics.jl_b200')
import numpy as np, spindyn as sd
L, nup = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (17, 8)
dt = np.complex128 if (len(sys.argv) > 3 and sys.argv[3] == "c128") else np.float64
m = sd.XXZChain(L, Jxy=0.7, Jz=1.3, hz=0.2, nup=nup)
print(m.info, flush=True)
psi = np.random.default_rng(1).standard_normal(m.dim).astype(dt)
out = np.empty_like(psi); sd.apply_rescaled_H_(out, psi, sd.apply_H_, m, 3.7, -0.4); print("done", out[:3])
