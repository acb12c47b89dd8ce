#!/usr/bin/env python
"""bench.py -- XXZ L=32 Sz=0 H.psi applies/s on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one H.psi (reference apply_H!, Hamiltonian.jl:211-273) over the
whole L=32, nup=16 sector (601 080 390 states, f64) with the counter-based
seeded psi.  `value` is device-resident throughput timed with CUDA events on the
library's own stream; `e2e` is the same call through the host-buffer entry
points (pinned H2D of psi, kernel, D2H of out inside the timed region);
`roofline` compares algorithmic bytes (16 B/state) with the measured HBM copy
peak; `cpu_baseline` times the CPU oracle port of the reference loop on a
bounded sample.  With N > 1 the same L=32 vector is sharded by rank range over
N GPUs (strong scaling): peer shards are read over NVLink, no data collective.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "spindynamics.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "XXZ L=32 Sz=0 H.psi applies/s"
UNIT = "applies/s"
SEED = 20261018


def comb(n, k):
    from math import comb as c
    return c(n, k)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML (nvidia_ml_py) is polled every ~2 ms when it
    initialises; otherwise `nvidia-smi --query-gpu` (one sample per ~100 ms) as in B200_PROFILING.md."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    NVML_BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t, self.source = index, [], threading.Event(), None, "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); int(reasons(h))      # probe once
            self._nvml = (pynvml, h, reasons, mx)
            self.source = "nvml"
        except Exception:
            self._nvml = None

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    nv, h, reasons, mx = self._nvml
                    r = int(reasons(h))
                    self.samples.append([str(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))), str(mx)] +
                                        ["Active" if r & self.NVML_BITS[n] else "Not Active" for n in self.NAMES])
                    self._stop.wait(0.002)
                    continue
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                    "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in o.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            if len(s) < 6:
                continue
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
            except ValueError:
                continue
            for n, v in zip(self.NAMES, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "source": self.source}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def mem_available_gb():
    try:
        with open("/proc/meminfo") as f:
            for ln in f:
                if ln.startswith("MemAvailable:"):
                    return int(ln.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_reference(L_sample, reps, L_target=32, ranked=True, warm=True):
    """Times the oracle's reference-faithful apply_H! (explicit states[], open-addressing table in place of the Dict
    with Julia's load factor, per-term loops, OpenMP over idx) on XXZ L_sample at ALL host threads: the team is set
    explicitly because torchrun exports OMP_NUM_THREADS=1 to its children.  1 warm-up apply, then `reps` timed ones.
    value = applies/s of the L_target problem: measured directly when L_sample == L_target (same_config), otherwise
    converted by the states*bonds ratio and labelled as such."""
    from oracle import oracle as orc
    nthr = host_threads()
    orc.lib().orc_set_num_threads(nthr)
    t_build = time.perf_counter()
    m = orc.XXZChain(L_sample, Jxy=1.0, Jz=1.0, hz=0.0, nup=L_sample // 2)
    n = len(m)
    psi = orc.fill_seeded(n, SEED)
    out = np.empty_like(psi)
    t_build = time.perf_counter() - t_build
    out[:] = 0.0                                                # page faults of out are not the algorithm's
    if warm:
        orc.apply_H_(out, psi, m)                               # warm-up (skipped at L >= 30: an apply takes ~1 min and the
    times = []                                                  # working set is 100x every cache, so there is nothing to warm)
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.apply_H_(out, psi, m)
        times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    work_sample = n * (L_sample - 1)
    work_target = comb(L_target, L_target // 2) * (L_target - 1)
    cores = int(orc.lib().orc_num_threads())
    same = L_sample == L_target
    res = {"value": (work_sample / t) / work_target, "unit": UNIT, "cores": cores, "kind": "port", "same_config": same,
           "sample": f"oracle/oracle.c orc_apply_H_f64 (C/OpenMP restatement of Hamiltonian.jl:211-273; Julia is not "
                     f"installed) on XXZ L={L_sample} nup={L_sample // 2} ({n} states), {1 if warm else 0} warm-up + {reps} timed applies, "
                     f"{t * 1e3:.1f} ms each on {cores} threads (basis + table built in {t_build:.0f} s, untimed)"
                     + ("" if same else f"; converted to L={L_target} by states*bonds"),
           "ms_per_sample_apply": t * 1e3, "host_threads_available": nthr}
    if ranked:
        # second, stronger CPU figure: same loop, Dict probe replaced by combinatorial ranking
        orc.apply_H_ranked_(out, psi, m)
        t0 = time.perf_counter()
        orc.apply_H_ranked_(out, psi, m)
        t_ranked = time.perf_counter() - t0
        res["ranked_value"] = (work_sample / t_ranked) / work_target
        res["ranked_note"] = (f"same loop with combinatorial ranking instead of the hash-map probe (not the reference's "
                              f"algorithm): {t_ranked * 1e3:.1f} ms per L={L_sample} apply")
    return res, times


def sampled_row_parity(model, out, scale, first, count, nrows, rank, cplx=False, periodic=False):
    """max over sampled rows r of this rank's shard of |out[r] - (H psi)[r]| / max(|(H psi)[r]|, scale), with (H psi)[r]
    from the oracle's row formula (oracle.c orc_row_seeded_f64 restates Hamiltonian.jl:223-269 for one state and
    regenerates psi from the counter).  Checker only: nothing here is timed."""
    from oracle import oracle as orc
    L, nup = model.L, model.nup
    if count == 0:
        return {"rows": 0, "max_rel_err": 0.0, "what": "sampled rows of H.psi vs oracle row formula"}
    rng = np.random.default_rng(1234 + rank)
    rows = np.unique(np.concatenate([rng.integers(first, first + count, max(nrows - 3, 1)),
                                     [first, first + count - 1, first + count // 2]])).astype(np.uint64)
    got, present = out.get(rows)
    assert present.all()
    hop = [(i, i + 1, 0.5) for i in range(1, L)]
    zz = [(i, i + 1, 1.0) for i in range(1, L)]
    if periodic:                                                 # SpinModel.jl:71-78
        hop.append((L, 1, 0.5))
        zz.append((L, 1, 1.0))
    worst = 0.0
    for r, g in zip(rows, got):
        state = int(model.unrank(int(r), 1)[0])
        ref = orc.row_seeded_f64((L, nup, hop, zz, np.zeros(L)), state, SEED, scale)
        if cplx:                                                 # H is real: the imaginary part is the row of the psi seeded with SEED + 1
            ref = complex(ref, orc.row_seeded_f64((L, nup, hop, zz, np.zeros(L)), state, SEED + 1, scale))
        worst = max(worst, abs(complex(g) - ref) / max(abs(ref), scale))
    return {"rows": int(len(rows)), "max_rel_err": worst, "what": "sampled rows of H.psi vs oracle row formula"}


def run_configs(args):
    """--configs-only: BASELINE.json configs 1 and 3 end to end through the public API on one GPU, with the CPU port of
    the same calls timed in the same run (config 3: a bounded CPU sample).  Prints {"configs": {...}}."""
    import spindyn as sd
    from oracle import oracle as orc
    orc.lib().orc_set_num_threads(host_threads())
    ctx = sd.Context(0)
    sd.set_default_context(ctx)
    res = {}
    # ---- config 1: XXZChain L=16 nup=8, groundstate(lanc_m=100) + lanczos_sqw(q = momenta, w = range(0,5,100), lanc_m=100, eta=0.05)
    L = 16
    rng = np.random.default_rng(SEED)
    wr = np.linspace(0.0, 5.0, 100)

    split = {}

    def config1(lib_, model):
        v0 = rng.standard_normal(len(model) if hasattr(model, "__len__") else model.dim)
        t_a = time.perf_counter()
        E0, psi = lib_.groundstate(model, lanc_m=100, v0=v0)
        t_b = time.perf_counter()
        S = lib_.lanczos_sqw(psi, model, lib_.momenta(model), wr, lanc_m=100, eta=0.05)
        split["groundstate_s"], split["lanczos_sqw_s"] = t_b - t_a, time.perf_counter() - t_b
        return E0, np.asarray(S)

    m = sd.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=L // 2, ctx=ctx)
    rng = np.random.default_rng(SEED)
    config1(sd, m)                                              # warm-up: module load, table uploads
    ctx.sync()
    l0 = ctx.launch_count()
    rng = np.random.default_rng(SEED)
    t0 = time.perf_counter()
    E_g, S_g = config1(sd, m)
    ctx.sync()
    t_gpu = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    split_gpu = dict(split)
    om = orc.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=L // 2)
    rng = np.random.default_rng(SEED)
    t0 = time.perf_counter()
    E_c, S_c = config1(orc, om)
    t_cpu = time.perf_counter() - t0
    res["config1"] = {"what": "XXZChain L=16 nup=8: groundstate(lanc_m=100) + lanczos_sqw(16 momenta, 100 frequencies, lanc_m=100, eta=0.05)",
                      "gpu_s": t_gpu, "cpu_port_s": t_cpu, "cpu_threads": host_threads(), "gpu_launches": int(launches),
                      "gpu_split": split_gpu, "cpu_split": dict(split),
                      "E0_gpu": float(E_g), "E0_cpu": float(E_c), "E0_abs_diff": abs(float(E_g) - float(E_c)),
                      "Sqw_rel_l2_diff": float(np.linalg.norm(S_g - S_c) / max(np.linalg.norm(S_c), 1e-300)),
                      "note": "S(q,w) from 100 unreorthogonalised Lanczos steps is ill-conditioned in the last Ritz values: see tests/test_gpu_sqw_tolerance.py for the tolerance"}
    del m
    # ---- config 3: XXZChain L=28 nup=14, KPM S(q,w) with 1024 Chebyshev moments (fused moment dots)
    L, M = 28, 1024
    m = sd.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=L // 2, ctx=ctx)
    psi0 = m.vector(np.float64).fill_seeded(SEED, 1.0 / np.sqrt(m.dim / 3.0))
    Eb = (-0.4432 * L - 0.5, (L - 1) / 4 + 0.25)
    a, b = (Eb[1] - Eb[0]) / (2 * 0.99), (Eb[1] + Eb[0]) / 2
    q = np.array([np.pi])
    w3 = np.linspace(0.0, 4.0, 200)
    sd.kpm_sqw(psi0, m, q, w3, a=a, b=b, kpm_m=8)               # warm-up
    ctx.sync()
    t0 = time.perf_counter()
    S3 = sd.kpm_sqw(psi0, m, q, w3, a=a, b=b, kpm_m=M)
    ctx.sync()
    t_gpu3 = time.perf_counter() - t0
    om = orc.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=L // 2)
    phi = orc.Sz_q_vector(om, orc.fill_seeded(len(om), SEED) * (1.0 / np.sqrt(len(om) / 3.0)), float(np.pi))
    Mc = 5
    t0 = time.perf_counter()
    orc.compute_chebyshev_moments(orc.apply_H_, phi, Mc, a, b, om)
    t_cpu3 = time.perf_counter() - t0
    res["config3"] = {"what": f"XXZChain L=28 nup=14: kpm_sqw, one momentum (q = pi), {M} Chebyshev moments, fused moment dots, c128",
                      "gpu_s": t_gpu3, "gpu_ms_per_moment": t_gpu3 * 1e3 / (M - 1),
                      "cpu_port_ms_per_moment": t_cpu3 * 1e3 / (Mc - 1), "cpu_sample": f"{Mc} moments of the same recurrence on {host_threads()} threads",
                      "cpu_port_s_extrapolated": t_cpu3 / (Mc - 1) * (M - 1), "sum_S_dw": float(np.sum(S3) * (w3[1] - w3[0]))}
    # ---- config 3 as BASELINE.json states it: q = momenta(model) (28 of them).  The q-batched path (one [state][q] multi-vector,
    # one fused kernel per moment for all momenta) against the per-momentum loop, on a reduced moment count so that the leg
    # stays within seconds; both scale linearly in the number of moments.
    try:
        qs = sd.momenta(m)
        Mq = 48
        sd.kpm_sqw(psi0, m, qs, w3, a=a, b=b, kpm_m=4, q_batch=True)             # warm-up
        ctx.sync()
        t0 = time.perf_counter()
        Sb = sd.kpm_sqw(psi0, m, qs, w3, a=a, b=b, kpm_m=Mq, q_batch=True)
        ctx.sync()
        t_b = time.perf_counter() - t0
        t0 = time.perf_counter()
        Sl = sd.kpm_sqw(psi0, m, qs[:4], w3, a=a, b=b, kpm_m=Mq, q_batch=False)
        ctx.sync()
        t_l = (time.perf_counter() - t0) * len(qs) / 4
        res["config3_all_momenta"] = {"what": f"XXZChain L=28 nup=14: kpm_sqw over the {len(qs)} momenta, {Mq} moments (timing sample; linear in the moments)",
                                      "batched_s": t_b, "batched_ms_per_moment_and_momentum": t_b * 1e3 / ((Mq - 1) * len(qs)),
                                      "q_loop_s_extrapolated_from_4_momenta": t_l, "q_loop_ms_per_moment_and_momentum": t_l * 1e3 / ((Mq - 1) * len(qs)),
                                      "batched_vs_loop_rel_l2": float(np.linalg.norm(Sb[:4] - Sl) / max(np.linalg.norm(Sl), 1e-300))}
    except Exception as exc:                                         # noqa: BLE001  (a newer leg must not cost the others)
        res["config3_all_momenta"] = {"error": repr(exc)[:300]}
    print(json.dumps({"configs": res}), flush=True)


def run_solve(args):
    """--solve-only: end-to-end solve on one GPU, everything device-resident; prints {"solve": {...}}."""
    import spindyn as sd
    ctx = sd.Context(0)
    sd.set_default_context(ctx)
    L = args.L
    model = sd.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=L // 2, ctx=ctx)
    v0 = model.vector(np.float64).fill_seeded(SEED + 1, 1.0)
    ctx.sync()
    t0 = time.perf_counter()
    l0 = ctx.launch_count()
    E0, gs = sd.lanczos_groundstate_lean(sd.apply_H_, model, lanc_m=args.solve_m, v0=v0, device=True)
    ctx.sync()
    wall = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    nrm = float(gs.norm())
    del gs
    # the same solve again: work vectors now come from the model's pool (no cudaMalloc / cudaFree of 4.8 GB buffers)
    t0 = time.perf_counter()
    E1, gs = sd.lanczos_groundstate_lean(sd.apply_H_, model, lanc_m=args.solve_m, v0=v0, device=True)
    ctx.sync()
    wall2 = time.perf_counter() - t0
    napply = 2 * args.solve_m - 1
    print(json.dumps({"solve": {"what": f"lanczos_groundstate_lean XXZ L={L} nup={L // 2}, lanc_m={args.solve_m} ({napply} H.psi with fused dot + "
                                        f"{2 * (args.solve_m - 1)} fused 3R+1W updates, device-resident scalars, 3 work vectors), seeded start vector",
                                "ms": wall * 1e3, "ms_second_call": wall2 * 1e3, "ms_per_lanczos_step": wall2 * 1e3 / napply,
                                "gpu_launches": int(launches), "E_ritz": float(E0), "E_ritz_second_call": float(E1), "ritz_norm": nrm}}), flush=True)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the bench's own configuration.  Julia is
    absent from this image, so this is the oracle port (kind "port"), on all host threads.  The L = 32 problem itself is
    run when the host has the memory (states 4.8 GB + table 17.2 GB + two vectors 9.6 GB); the number of timed applies
    does not depend on --steps (a CPU apply takes about a minute): 2 timed applies."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    L_t = args.L
    need_gb = {32: 40.0, 30: 12.0}
    avail = mem_available_gb()
    L_s = L_t
    if args.cpu_L_ref:
        L_s = args.cpu_L_ref
    else:
        for cand in (L_t, 30, 28):
            if cand <= L_t and avail >= need_gb.get(cand, 4.0):
                L_s = cand
                break
    reps = 2 if L_s >= 30 else 5
    res, times = cpu_reference(L_s, reps, L_target=L_t, ranked=False, warm=L_s < 30)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": reps, "warmup": 1 if L_s < 30 else 0, "ms_per_step": 1e3 / res["value"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (counter-based seeded psi)",
            "config": {"workload": f"XXZChain L={L_t} nup={L_t // 2} open Jxy=Jz=1 hz=0, f64 H.psi on the host CPU"
                                   + ("" if res["same_config"] else f" (sample L={L_s}, converted; host MemAvailable {avail:.0f} GB)"),
                       "requested_steps": args.steps, "requested_warmup": args.warmup,
                       "note": "a CPU apply takes about a minute: 2 timed applies whatever --steps says"},
            "same_config": res["same_config"],
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--L", type=int, default=32, help="chain length (default: the metric's L=32)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "c128"])
    ap.add_argument("--boundary", default="open", choices=["open", "periodic"], help="periodic: the wrap-bond pass + block kernel (not the headline metric)")
    ap.add_argument("--cpu-L", type=int, default=28, help="chain length of the bounded cpu_baseline sample of the GPU arm")
    ap.add_argument("--cpu-L-ref", type=int, default=0, help="--impl reference: force the chain length (default: L itself if the host has the memory)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled-row check of H.psi against the oracle")
    ap.add_argument("--parity-rows", type=int, default=256, help="sampled rows per rank")
    ap.add_argument("--no-solve", action="store_true", help="skip the end-to-end Lanczos solve leg")
    ap.add_argument("--solve-m", type=int, default=30)
    ap.add_argument("--solve-only", action="store_true", help="(internal) run only the solve leg and print {\"solve\": ...}")
    ap.add_argument("--configs-only", action="store_true", help="(internal) BASELINE.json configs 1 and 3 end to end, GPU vs CPU port; prints {\"configs\": ...}")
    ap.add_argument("--no-configs", action="store_true", help="skip the config 1 / config 3 legs")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--path", default=None, choices=[None, "block", "tiled", "generic"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.solve_only:
        return run_solve(args)
    if args.configs_only:
        return run_configs(args)

    import spindyn as sd
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        ctx = sd.Context.from_torch_distributed(local_rank)
    else:
        if sd.device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: libspindyn_cuda has no CPU fallback")
        ctx = sd.Context(0)
    sd.set_default_context(ctx)

    L, nup = args.L, args.L // 2
    dtype = np.float64 if args.dtype == "f64" else np.complex128
    esz = 8 if args.dtype == "f64" else 16
    model = sd.XXZChain(L, Jxy=1.0, Jz=1.0, hz=0.0, nup=nup, boundary=args.boundary, ctx=ctx)
    if args.path:
        model.set_path(args.path)
    N = model.dim
    first, count = model.local_range
    psi_scale = 1.0 / np.sqrt(N / 3.0)
    psi = model.vector(dtype).fill_seeded(SEED, psi_scale)                   # ||psi|| ~ 1
    out = model.vector(dtype)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(max(args.warmup, 0)):
        sd.apply_H_(out, psi, model)
    barrier()
    l0 = ctx.launch_count()
    with ClockSampler(local_rank) as clk:
        ctx.timer_start()
        for _ in range(args.steps):
            sd.apply_H_(out, psi, model)
        ms_total = ctx.timer_stop()
        barrier()
    launches = ctx.launch_count() - l0
    if dist is not None:
        import torch
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # parity of what was just timed: sampled rows of out = H psi on every rank against the oracle's row formula (the
    # oracle regenerates the counter-based psi, so this works at sizes no host can hold).  Outside the timed region.
    parity = None
    if not args.no_parity:
        parity = sampled_row_parity(model, out, psi_scale, first, count, args.parity_rows, rank, args.dtype == "c128",
                                    args.boundary == "periodic")
        if dist is not None:
            import torch
            t = torch.tensor([parity["max_rel_err"]], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r = torch.tensor([parity["rows"]], device="cuda", dtype=torch.int64)
            dist.all_reduce(r, op=dist.ReduceOp.SUM)
            parity = {**parity, "max_rel_err": float(t.item()), "rows": int(r.item()), "ranks": world}
        parity["ok"] = bool(parity["max_rel_err"] <= 1e-13)

    # end to end through the host-buffer entry points (pinned host memory)
    e2e = None
    checksum = None
    if not args.no_e2e and 2 * count * esz > (10 << 30):
        e2e = {"skipped": f"2 x {count * esz / 1e9:.1f} GB of pinned host memory per rank"}
    elif not args.no_e2e:
        hin = sd.PinnedBuffer(count, dtype)
        hout = sd.PinnedBuffer(count, dtype)
        hin.array[:] = 0.0
        psi.to_host(hin.array)
        lib, check = sd.lib(), sd._lib.check
        # two device-side (psi, out) pairs: step s + 1's upload runs while step s's result is still on its way to the
        # host (the library's copy engine keeps both PCIe directions busy); every step still moves its own input and
        # its own result inside the timed region, and the stopwatch stops after the last byte has landed
        pairs = [(psi, out), (model.vector(dtype), model.vector(dtype))]

        def e2e_step(s):
            p, o = pairs[s % 2]
            check(lib.sd_vec_upload_async(p._h, hin._p))
            check(lib.sd_apply_H(model._h, o._h, p._h))
            check(lib.sd_vec_download_async(o._h, hout._p))

        e2e_step(0)
        e2e_step(1)
        barrier()
        t0 = time.perf_counter()
        ctx.timer_start()
        for s in range(args.e2e_steps):
            e2e_step(s)
        ms_e2e = ctx.timer_stop()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms_e2e = max(ms_e2e, 0.0)
        if dist is not None:
            import torch
            t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        e2e = {"value": 1e3 / (ms_e2e / args.e2e_steps), "unit": UNIT,
               "h2d_bytes_per_step": int(N * esz), "d2h_bytes_per_step": int(N * esz),
               "steps": args.e2e_steps, "ms_per_step": ms_e2e / args.e2e_steps, "wall_ms_per_step": wall / args.e2e_steps,
               "note": "pinned host psi -> HBM, sd_apply_H, out -> pinned host, all inside the timed region; two steps in flight "
                       "(step s + 1's upload overlaps step s's download: both PCIe directions busy), timed until the last download has landed"}
        checksum = float(np.sum(hout.array[:min(count, 1 << 20)]).real)
        del pairs
        hin.free()
        hout.free()
    else:
        checksum = None

    # end-to-end solve on the same model (north star: "end-to-end solve time"): 30 steps of the memory-lean Lanczos
    # ground state (three work vectors; the reference's N x m basis does not fit one GPU beyond m = 33 at L = 32).
    # Runs in a child process under a timeout after everything the contract needs has been measured, so that a
    # failure of this (newer) code path can only cost its own entry, never the bench line.
    solve = None
    if not args.no_solve and not args.no_e2e and args.dtype == "f64" and world == 1:
        import subprocess
        del psi, out
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--solve-only", "--L", str(L), "--solve-m", str(args.solve_m)],
                                capture_output=True, text=True, timeout=240)
            last = [ln for ln in cp.stdout.splitlines() if ln.startswith("{")]
            solve = json.loads(last[-1])["solve"] if (cp.returncode == 0 and last) else {"error": (cp.stderr or cp.stdout)[-300:]}
        except Exception as exc:                                   # noqa: BLE001
            solve = {"error": repr(exc)[:300]}
    configs = None
    if not args.no_configs and not args.no_cpu and not args.no_e2e and args.dtype == "f64" and world == 1:
        import subprocess
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--configs-only"], capture_output=True, text=True, timeout=300)
            last = [ln for ln in cp.stdout.splitlines() if ln.startswith("{")]
            configs = json.loads(last[-1])["configs"] if (cp.returncode == 0 and last) else {"error": (cp.stderr or cp.stdout)[-300:]}
        except Exception as exc:                                   # noqa: BLE001
            configs = {"error": repr(exc)[:300]}

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = 2 * esz * N                              # read psi once + write out once
        # the apply kernel is the only kernel in a step; per-launch time = step time (per GPU: its shard)
        achieved = alg_bytes / world / (ms_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        knobs = {k: v for k, v in os.environ.items() if k.startswith(("SD_BLK", "SD_HALO", "SD_SHARD")) or k in ("SD_FAR_MB", "SD_FORCE_GENERIC")}
        kname = {"block": "sd_blkl_apply_kernel", "tiled": "sd_tile_apply_kernel"}.get(model.info["kernel_path"], "sd_generic_apply_kernel")
        if os.path.exists(tpath) and world == 1 and not knobs:   # the ncu capture is of the default single-GPU launch
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get(f"apply_L{L}_{args.dtype}_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "kernel": kname,
                    "algorithmic_bytes_per_launch": alg_bytes // world,
                    "note": "16 B/state f64 (32 c128): one read of psi + one write of out; index math is on the fly"}
        cpu = None
        if not args.no_cpu and world == 1:
            cpu, _ = cpu_reference(args.cpu_L, 4, L_target=L)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic (counter-based seeded psi, splitmix64)",
                "config": {"workload": f"XXZChain L={L} nup={nup} {args.boundary} Jxy=Jz=1 hz=0, {args.dtype} H.psi, N={N} states "
                                       f"({N * esz / 1e9:.2f} GB per vector)",
                           "l2": "inputs >> 126 MB L2, no flush needed" if N * esz > 1e9 else "WARNING: fits L2",
                           "sharding": f"{world} contiguous rank ranges, NVLink peer reads" if world > 1 else "single GPU",
                           "kernel_path": model.info["kernel_path"], "tile_sites": model.info["tile_sites"],
                           **({"env_knobs": knobs} if knobs else {})},
                "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
                "parity": parity, "cpu_baseline": cpu, "checksum": checksum, "solve": solve, "configs": configs}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
