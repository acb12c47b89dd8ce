import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/spindynamics.jl_b200')
import numpy as np, spindyn as sd
m = sd.XXZChain(18, Jxy=0.7, Jz=1.3, hz=0.2, nup=9)
psi = np.random.default_rng(1).standard_normal(m.dim)
out = np.empty_like(psi); sd.apply_H_(out, psi, m); print("done", out[:3])
