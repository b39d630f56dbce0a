#!/usr/bin/env python
"""Benchmark of the GPJax hot path on B200 (driver contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|exact|sgpr]

Headline (`value`): exact-GP ``conjugate_mll`` value+gradient evaluations per second at N=50,000, D=8,
float64 (BASELINE.json metric; the configuration the metric is quoted on fits one GPU).  The exact
path does not shard ("replicas only"), so ``--gpus N`` runs N independent replicas (weak
scaling) -- and ALSO runs the path that does shard, SGPR ``collapsed_elbo`` at N=10M / M=2048 row-sharded
over the ranks with an NCCL all-reduce, reported under the ``sgpr`` key (points/s, strong scaling).
``--workload sgpr`` makes the SGPR number the main ``value`` instead.

Both paths run their O(N^3) / O(N M^2) products as exact int8 digit-plane (Ozaki) products on tcgen05 by default (DESIGN
section 12); the roofline object then describes that kernel, and `fp64_dmma_path_ms_per_step` / `remaining_dmma_gemms` the
FP64 DMMA path (GPB_OZAKI=0 in the environment benches it alone).

Timing: W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the
launching stream, MAX over ranks.  L2: every step streams a 20 GB (exact) / multi-GB (SGPR) working set,
far beyond the 126 MB L2, so no explicit flush is needed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "exact-GP MLL+grad evals/s @N=50k fp64; SGPR ELBO points/s at 1/2/4/8 GPU"
NOMINAL_8BIT_TOPS = 4500.0  # dense fp8 / int8-class tensor peak of B200 (B200_PROFILING.md: 4.5 PFLOP/s dense fp8)
NOMINAL_FP64_TFLOPS = 128 * 148 * 1.965e9 / 1e12  # 128 flop/clk/SM x 148 SMs x 1.965 GHz = 37.2


def measured_peaks():
    """Driver-written MEASURED_PEAKS.json (HBM copy GB/s, bf16 TF/s); it holds no FP64 figure."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return {"hbm_gbs": d.get("hbm_gbs"), "bf16_tflops": d.get("bf16_tflops"),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained")}
    except Exception:
        return None


def env_int(name, default):
    return int(os.environ.get(name, default))


# ----------------------------------------------------------------------------------------------
# synthetic data: X ~ U(-2,2)^{N x D}, y = sin(x0) + 0.1 eps, NumPy PCG64
# ----------------------------------------------------------------------------------------------
def synth(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-2.0, 2.0, (n, d))
    y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((n, 1))
    return X, y


SYNTH_CHUNK = 65536


def synth_rows(lo, hi, d, seed):
    """Rows [lo, hi) of ONE global synthetic data set, independent of how the rows are sharded: chunk c (65,536 rows) is
    drawn from its own PCG64 stream seeded (seed, c), so every world size sees the same N rows and the ELBO printed at
    1 / 2 / 4 / 8 GPUs can be cross-checked."""
    Xs, ys = [], []
    for c in range(lo // SYNTH_CHUNK, (hi - 1) // SYNTH_CHUNK + 1):
        rng = np.random.default_rng([seed, c])
        X = rng.uniform(-2.0, 2.0, (SYNTH_CHUNK, d))
        y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((SYNTH_CHUNK, 1))
        a, b = max(lo, c * SYNTH_CHUNK) - c * SYNTH_CHUNK, min(hi, (c + 1) * SYNTH_CHUNK) - c * SYNTH_CHUNK
        Xs.append(X[a:b]), ys.append(y[a:b])
    return np.concatenate(Xs), np.concatenate(ys)


HYPER = dict(variance=1.0, obs_stddev=0.3, mean_const=0.0, jitter=1e-6)


def ell_ard(d):
    return np.linspace(0.8, 1.6, d)


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background during the timed region)
# ----------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), smax.append(float(f[2])), power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port in the reference's operation order (jax absent)
# ----------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_exact_sample(n_target, d, budget_s=12.0):
    """One MLL+grad evaluation of the reference formulation on a bounded sample, N^3-extrapolated."""
    import oracle

    def one(ns):
        X, y = synth(ns, d, 50)
        t0 = time.perf_counter()
        oracle.reference_cpu_mll_value_and_grad("rbf", X, y, ell_ard(d), HYPER["variance"], HYPER["obs_stddev"],
                                                HYPER["mean_const"], HYPER["jitter"])
        return time.perf_counter() - t0

    t_small = one(1500)
    ns = int(min(8000, max(2000, 1500 * (budget_s / max(t_small, 1e-3)) ** (1 / 3))))
    ns = (ns // 500) * 500
    t = one(ns)
    return {"n_sample": ns, "seconds": t, "evals_per_s_at_target": 1.0 / (t * (n_target / ns) ** 3)}


def cpu_sgpr_sample(m, d, budget_rows=40000):
    import oracle

    X, y = synth(budget_rows, d, 4)
    Z = synth(m, d, 5)[0]
    t0 = time.perf_counter()
    oracle.reference_cpu_elbo_value_and_grad("rbf", X, y, Z, ell_ard(d), HYPER["variance"], HYPER["obs_stddev"],
                                             HYPER["mean_const"], HYPER["jitter"], block=8192)
    t = time.perf_counter() - t0
    return {"rows": budget_rows, "seconds": t, "points_per_s": budget_rows / t}


def run_reference(args):
    """--impl reference: the reference's own CPU formulation (oracle port; jax[cpu] cannot be installed
    here: no jax/jaxlib wheels in /opt/wheelhouse, no network) on the box's host cores."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    n, d = env_int("GPB_BENCH_N", 50000), 8
    cores = cpu_threads()
    vals, ts = [], []
    samp = None
    for i in range(args.warmup + args.steps):
        samp = cpu_exact_sample(n, d, budget_s=float(os.environ.get("GPB_REF_BUDGET_S", "10")))
        if i >= args.warmup:
            vals.append(samp["evals_per_s_at_target"])
            ts.append(samp["seconds"])
    v = float(np.mean(vals))
    sample = (f"one MLL+grad (LU slogdet + LU solve + full inverse, reference operation order) at "
              f"N={samp['n_sample']}, D={d} of the same synthetic data, {np.mean(ts):.2f} s per eval, "
              f"N^3-extrapolated to N={n}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"exact_gp_conjugate_mll_value_and_grad_N{n}_D{d}_RBF_ARD", "N": n, "D": d,
                       "timed_sample_N": samp["n_sample"],
                       "extrapolation": f"each step times the reference formulation at N={samp["n_sample"]} and scales by (N/N_sample)^3 "
                                        f"to N={n}: the full size would take ~80 min per step on these host cores"},
            "cpu_baseline": {"value": v, "unit": "evals/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Dist:
    def __init__(self, n_gpus):
        import torch

        self.torch = torch
        self.world = env_int("WORLD_SIZE", 1)
        self.rank = env_int("RANK", 0)
        self.local = env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl ours) needs a CUDA device: gpjax_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        else:
            self.dist = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())


def timed(D: Dist, step, warmup, steps):
    torch = D.torch
    for _ in range(warmup):
        step()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    D.barrier()
    return D.max_over_ranks(t)


def ncu_capture(rel_path):
    """`traffic` of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, parsed from the committed
    summary of an `ncu --set full` capture of this bench command (scripts/parse_ncu.py writes it), next to the algorithmic bytes
    of that launch's shape.  Absent file -> traffic null."""
    try:
        with open(os.path.join(ROOT, rel_path)) as f:
            d = json.load(f)
        return {"traffic": d["dram_bytes_read"] + d["dram_bytes_write"], "traffic_unit": "bytes per launch (the captured launch)",
                "algorithmic_bytes": d["algorithmic_bytes"], "traffic_source": rel_path, "traffic_launch": d.get("launch"),
                "tensor_pipe_active_pct_ncu": d.get("tensor_pipe_active_pct")}
    except Exception:
        return {"traffic": None}


def measure_cublas_dgemm(D: Dist):
    torch = D.torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=D.dev)
    b = torch.randn(n, n, dtype=torch.float64, device=D.dev)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    return 3 * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12


def measure_int8_ceiling(D: Dist, n=8192, sustain_s=2.0):
    """Int8 tensor ceiling of THIS box, measured in THIS process (MEASURED_PEAKS.json has no int8 entry): cuBLASLt IGEMM through
    torch._int_mm at n^3, burst = best of 10 single launches, sustained = back-to-back launches for >= sustain_s seconds
    (the figure a kernel timed inside a long step is compared with), SM clock sampled during the sustained loop."""
    torch = D.torch
    g = torch.Generator(device=D.dev).manual_seed(1)
    a = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=D.dev, generator=g)
    b = torch.randint(-128, 128, (n, n), dtype=torch.int8, device=D.dev, generator=g).t()  # column-major B, as cuBLASLt wants
    for _ in range(3):
        torch._int_mm(a, b)
    torch.cuda.synchronize()
    ops_per = 2.0 * float(n) ** 3
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, ops_per / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    per = ops_per / (best * 1e12)
    reps = max(10, int(sustain_s / per))
    clocks = Clocks(D.local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch._int_mm(a, b)
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    sustained = reps * ops_per / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return {"how": f"torch._int_mm (cuBLASLt IGEMM) {n}^3: best of 10 launches (burst), {reps} back-to-back launches (sustained)",
            "burst_tops": best, "sustained_tops": sustained, "sustained_seconds": e0.elapsed_time(e1) * 1e-3,
            "sm_mhz_sustained": clk.get("sm_mhz"), "power_w_max": clk.get("power_w_max"), "reasons": clk.get("reasons")}


def int8_roofline_peak(D: Dist):
    """Peak of the int8 roofline of a kernel that shares its step with FP64 kernels (SGPR / SVGP: the int8 kernel is 60-65 % of
    the step, so the 1 kW cap that throttles back-to-back IGEMMs does not bind it): the BURST cuBLASLt IGEMM rate measured live in
    this process; the sustained (power-capped) figure and 2 x bf16-sustained are reported next to it."""
    i8 = measure_int8_ceiling(D)
    mp = measured_peaks() or {}
    bf16x2 = 2.0 * float(mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops") or 1414.5)
    src = ("measured live in this process: torch._int_mm (cuBLASLt IGEMM) 8192^3, best single launch (burst: this kernel alternates with "
           "FP64 kernels, the power cap that limits back-to-back IGEMMs does not bind it); unit is int8 Top/s (2 x MAC).  The best-of-10 "
           "figure moves by +-3 % between processes, so frac ~ 1 reads as parity with the library kernel on long-K products, not as a "
           "hardware limit: frac_of_nominal_dense_8bit_peak is the fraction of the 4.5 Pop/s dense 8-bit tensor peak")
    return i8, bf16x2, src


def measure_hbm_kernels(D: Dist, n, d, X, ell, var):
    """HBM rooflines of the epilogue / solve kernels the north-star names, against MEASURED_PEAKS.json's copy bandwidth:
    gram_kernel (lower triangle of Sigma written once: 8 B per entry; ~40 FP64 instructions per entry make it FP64-issue bound
    before it is HBM bound) and the GEMV of the blocked triangular solve (8 B per factor entry read once)."""
    torch = D.torch
    from gpjax_b200 import ops

    mp = measured_peaks() or {}
    peak = float(mp.get("hbm_gbs") or 6456.2)
    st = ops._mll_state(n, d, X.device)  # the step's own N x N buffer (holds Sigma^-1 / L after the last backward)
    def ev_ms(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    out = {"peak": peak, "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth)"}
    with torch.no_grad():
        t_gram = ev_ms(lambda: ops.gram_forward(0, X, X, ell.detach(), var.detach(), diag_add=1e-6 + 0.09, lower_only=True, out=st.sigma))
        live = n * (n + 1) / 2 + n * 64  # entries of the 64 x 128 tiles that touch the lower triangle (upper bound: + one tile row)
        out["gram_kernel"] = {"ms": t_gram, "algorithmic_bytes": 8.0 * n * (n + 1) / 2, "achieved": 8.0 * live / t_gram / 1e6,
                              "frac": 8.0 * live / t_gram / 1e6 / peak,
                              "note": "FP64-issue bound (exp + direct differences, ~40 FP64 instructions per entry) at this D"}
        # blocked triangular solve L w = d (gemv_n / gemv_t kernels + the stored diagonal-block inverses): reads the factor once
        ws = ops.FactorWorkspace(n, d, potri=False, device=X.device)
        ops.potrf_lower_(st.sigma, ws, zero_upper=False)
        v = torch.ones(n, dtype=torch.float64, device=X.device)
        for trans, name in ((False, "trsv_forward"), (True, "trsv_transposed")):
            t_s = ev_ms(lambda: ops.trsv_lower_(st.sigma, v.clone(), ws, trans=trans))
            out[name] = {"ms": t_s, "algorithmic_bytes": 8.0 * n * (n + 1) / 2, "achieved": 8.0 * n * (n + 1) / 2 / t_s / 1e6,
                         "frac": 8.0 * n * (n + 1) / 2 / t_s / 1e6 / peak,
                         "note": "gemv_n / gemv_t over the off-diagonal panels + one GEMV per stored diagonal-block inverse"}
        del ws
    return out


def bench_exact(D: Dist, args):
    torch = D.torch
    import gpjax_b200 as gpx
    from gpjax_b200 import ops
    from gpjax_b200._lib import lib
    from gpjax_b200.parameters import NonNegativeReal, PositiveReal, Real

    n, d = env_int("GPB_BENCH_N", 50000), 8
    Xn, yn = synth(n, d, 50 + D.rank)
    Xh, yh = torch.from_numpy(Xn).pin_memory(), torch.from_numpy(yn).pin_memory()
    X, y = Xh.to(D.dev), yh.to(D.dev)
    mk = lambda v: torch.as_tensor(np.asarray(v, np.float64), device=D.dev).requires_grad_(True)
    ell, var, sn, c = mk(ell_ard(d)), mk(HYPER["variance"]), mk(HYPER["obs_stddev"]), mk(HYPER["mean_const"])
    params = (ell, var, sn, c)

    def step():
        for p in params:
            p.grad = None
        v = ops.conjugate_mll_fused(0, X, y, ell, var, sn, c, HYPER["jitter"])
        v.backward()

    L = lib()
    clocks = Clocks(D.local)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    L.gpb_profile_reset(1)
    clocks.start()
    t = timed(D, step, 0, args.steps)
    clk = clocks.stop()
    import ctypes as C

    gemm_ms, gemm_n, all_n = C.c_double(), C.c_int64(), C.c_int64()
    oz_ms, oz_n, oz_ops = C.c_double(), C.c_int64(), C.c_double()
    L.gpb_profile_read(C.byref(gemm_ms), C.byref(gemm_n), C.byref(all_n))
    L.gpb_profile_read_ozaki(C.byref(oz_ms), C.byref(oz_n), C.byref(oz_ops))
    L.gpb_profile_reset(0)
    value = D.world * args.steps / t
    flops_per_eval = float(n) ** 3  # N^3/3 potrf + 2N^3/3 potri
    mode = int(L.gpb_get_ozaki_slices())  # -1: auto (device-side conditioning guard), 0: DMMA only, 4..7: forced
    planes = (int(L.gpb_ozaki_auto_planes(n, HYPER["variance"], HYPER["obs_stddev"], HYPER["jitter"])) if mode == -1 else mode)
    dmma = {"kernel": "gemm_f64_kernel (FP64 DMMA.8x8x4)", "launches_per_step": gemm_n.value / args.steps,
            "time_over_step_time": gemm_ms.value * 1e-3 / t}
    if planes and oz_n.value > 0:
        # Dominant kernel: ozaki_i8_kernel_w4 (tcgen05.mma.cta_group::2.kind::i8).  `achieved` = algorithmic int8 operations of
        # its launches (live output entries x K x digit pairs x 2) / summed launch durations (CUDA events on the launching
        # stream).  Digit pairs of s planes: s(s+1)/2, plus the equal-plane pair (s/2, s/2) when s is even (csrc/ozaki_i8.cu:
        # oz_has_diag).  The launcher counts with the 7 planes the digit buffers are laid out for (28 pairs); the device-side
        # guard used `planes` of them.
        pairs = lambda q: q * (q + 1) // 2 + (1 if (q >= 2 and q % 2 == 0) else 0)
        oz_ops_used = oz_ops.value * pairs(planes) / 28.0
        achieved = oz_ops_used / (oz_ms.value * 1e-3) / 1e12
        # MEASURED_PEAKS.json has no int8 entry: the ceiling is measured live in this process (cuBLASLt IGEMM, torch._int_mm):
        # `peak` = the SUSTAINED figure (the kernel is timed inside a seconds-long step under the 1 kW cap), burst alongside.
        i8 = measure_int8_ceiling(D)
        mp = measured_peaks() or {}
        bf16 = float(mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops") or 1414.5)
        ncu = ncu_capture("profiles/r02_ozaki_w4_nb2048_ncu.json" if L.gpb_block_size_for(n) == 2048 else "profiles/r02_ozaki_w4_ncu.json")
        roof = {"bound": "tensor",
                "kernel": "ozaki_i8_kernel_w4 (tcgen05.mma.cta_group::2.kind::i8, M256 x N256 per CTA pair = two digit-pair orders side by side, "
                          "int8 x int8 -> int32 in TMEM, int64 fixed-point recombination, red.global.add.f64 write-out)",
                "achieved": achieved, "peak": i8["sustained_tops"], "unit": "TFLOP/s", "frac": achieved / i8["sustained_tops"],
                "peak_source": "measured live in this process: torch._int_mm (cuBLASLt IGEMM) 8192^3 back to back for >= 2 s "
                               "(sustained, power-capped); unit is int8 Top/s (2 x MAC)",
                "int8_ceiling_measured": i8, "frac_of_int8_burst": achieved / i8["burst_tops"],
                "frac_of_nominal_dense_8bit_peak": achieved / NOMINAL_8BIT_TOPS,
                "note": "frac > 1 against the sustained library figure means this kernel sustains more int8 Top/s than cuBLASLt's IGEMM "
                        "under the same 1 kW power cap (sw_power_cap is the limiter of both); the pipe-limited comparison is "
                        "frac_of_int8_burst",
                "frac_of_2x_bf16_sustained": achieved / (2.0 * bf16),
                "digit_planes": planes, "digit_bits_per_plane": 8, "digit_pair_products": pairs(planes), "digit_plane_mode": "auto (device-side conditioning guard)" if mode == -1 else "forced",
                "int8_ops_per_eval": oz_ops_used / args.steps,
                "launches_per_step": oz_n.value / args.steps, "time_over_step_time": oz_ms.value * 1e-3 / t,
                "algorithmic_flop_per_eval": flops_per_eval,
                "whole_step_fp64_equivalent_tflops": flops_per_eval * args.steps / t / 1e12,
                "whole_step_fp64_equivalent_over_dmma_peak_NOT_a_roofline_fraction": flops_per_eval * args.steps / t / 1e12 / NOMINAL_FP64_TFLOPS,
                "fp64_dmma_peak": NOMINAL_FP64_TFLOPS, "measured_peaks_json": mp, "remaining_dmma_gemms": dmma}
        roof.update(ncu)
        # the same step with every update on the FP64 DMMA pipe (GPB_OZAKI=0), timed live for comparison
        ops.set_ozaki_slices(0)
        t_dmma = timed(D, step, 1, 1)
        ops.set_ozaki_slices(mode)
        roof["fp64_dmma_path_ms_per_step"] = 1e3 * t_dmma
    else:
        achieved = flops_per_eval * args.steps / (gemm_ms.value * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": dmma["kernel"], "achieved": achieved,
                "peak": NOMINAL_FP64_TFLOPS, "unit": "TFLOP/s", "frac": achieved / NOMINAL_FP64_TFLOPS,
                "peak_source": "nominal FP64 DMMA peak 128 flop/clk/SM x 148 SM x 1.965 GHz (MEASURED_PEAKS.json has no "
                               "FP64 figure); cuBLAS DGEMM measured live alongside",
                "peak_cublas_dgemm": measure_cublas_dgemm(D), "measured_peaks_json": measured_peaks(),
                "algorithmic_flop_per_eval": flops_per_eval,
                "gemm_launches_per_step": gemm_n.value / args.steps,
                # sum of GEMM launch durations / step time; can reach ~1.0 because the lookahead GEMMs on the side stream
                # overlap the trailing update on the main stream
                "gemm_time_over_step_time": gemm_ms.value * 1e-3 / t,
                "whole_step_tflops": flops_per_eval * args.steps / t / 1e12,
                # ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the 1846 GEMM launches of ONE evaluation at
                # N=50k with the 1024 block (profiles/r01_gemm_traffic_exact50k_nb1024.md: 1147 GB read + 499 GB written);
                # algorithmic = read + write of every C tile touched by the rank-NB updates of the three N^3/3 phases
                "traffic": 1.646e12 if (n == 50000 and L.gpb_block_size_for(n) == 1024) else None,
                "traffic_unit": "bytes per evaluation (all GEMM launches)",
                "algorithmic_bytes": 3 * 16 * float(n) ** 3 / (6 * L.gpb_block_size_for(n))}

    # ---- bandwidth-class kernels of the path, each timed alone on the benched shape (CUDA events on the launching stream) -------
    roof["roofline_hbm"] = measure_hbm_kernels(D, n, d, X, ell, var)

    # ---- e2e: the public API with HOST buffers (pinned), H2D + D2H inside the timed region -------------------
    prior = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(HYPER["mean_const"])),
                          kernel=gpx.kernels.RBF(lengthscale=PositiveReal(ell_ard(d)),
                                                 variance=NonNegativeReal(HYPER["variance"])), jitter=HYPER["jitter"])
    post = prior * gpx.likelihoods.Gaussian(num_datapoints=n, obs_stddev=NonNegativeReal(HYPER["obs_stddev"]))
    leaves = [p for _, p in post.named_parameters()]
    host_out = {}

    def e2e_step():
        data = gpx.Dataset(X=Xh.to(D.dev, non_blocking=True), y=yh.to(D.dev, non_blocking=True))
        vals = [p.value.detach().requires_grad_(True) for p in leaves]
        for p, v in zip(leaves, vals):
            p.value = v
        loss = -gpx.objectives.conjugate_mll(post, data)
        grads = torch.autograd.grad(loss, vals)
        host_out["loss"] = loss.item()
        host_out["grads"] = [g.cpu() for g in grads]

    e2e_steps = max(1, min(args.steps, 3))
    t_e2e = timed(D, e2e_step, 1, e2e_steps)
    e2e = {"value": D.world * e2e_steps / t_e2e, "unit": "evals/s", "h2d_bytes_per_step": int(n * d * 8 + n * 8),
           "d2h_bytes_per_step": int(8 * (1 + d + 3)), "steps": e2e_steps,
           "api": "gpx.objectives.conjugate_mll(posterior, Dataset) + autograd, pinned host X/y copied every step"}
    ops.release_buffers()
    torch.cuda.empty_cache()
    return dict(n=n, d=d, t=t, value=value, roofline=roof, clocks=clk, e2e=e2e, gpu_launches=int(all_n.value),
                workload=f"exact_gp_conjugate_mll_value_and_grad_N{n}_D{d}_RBF_ARD")


def bench_sgpr(D: Dist, args, steps=None, warmup=None):
    torch = D.torch
    import gpjax_b200 as gpx
    from gpjax_b200 import sgpr_ops
    from gpjax_b200._lib import lib
    from gpjax_b200.parameters import NonNegativeReal, PositiveReal, Real

    n_total, m, d = env_int("GPB_BENCH_SGPR_N", 10_000_000), env_int("GPB_BENCH_SGPR_M", 2048), 8
    block = env_int("GPB_BENCH_SGPR_BLOCK", 65536)
    steps = steps or max(5, args.steps)
    warmup = warmup if warmup is not None else max(3, min(args.warmup, 3))
    lo, hi = D.rank * n_total // D.world, (D.rank + 1) * n_total // D.world
    Xn, yn = synth_rows(lo, hi, d, 4)  # rank-independent: the same N rows at every world size
    Xh, yh = torch.from_numpy(Xn).pin_memory(), torch.from_numpy(yn).pin_memory()
    X, y = Xh.to(D.dev), yh.to(D.dev)
    Zn = synth(m, d, 5)[0]
    mk = lambda v: torch.as_tensor(np.asarray(v, np.float64), device=D.dev).requires_grad_(True)
    Z, ell, var, sn, c = mk(Zn), mk(ell_ard(d)), mk(HYPER["variance"]), mk(HYPER["obs_stddev"]), mk(HYPER["mean_const"])
    params = (Z, ell, var, sn, c)
    # exchange step through the C ABI (gpb_allreduce_f64 on the launching stream); torch.distributed only ships the unique id
    native = sgpr_ops.init_native_collective() if (D.world > 1 and os.environ.get("GPB_NATIVE_NCCL", "1") != "0") else False
    last = {}

    def step():
        for p in params:
            p.grad = None
        v = sgpr_ops.collapsed_elbo_fused(0, X, y, Z, ell, var, sn, c, HYPER["jitter"], block)
        v.backward()
        last["v"] = v.detach()

    L = lib()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    L.gpb_profile_reset(1)
    t = timed(D, step, 0, steps)
    import ctypes as C

    gemm_ms, gemm_n, all_n = C.c_double(), C.c_int64(), C.c_int64()
    oz_ms, oz_n, oz_ops = C.c_double(), C.c_int64(), C.c_double()
    L.gpb_profile_read(C.byref(gemm_ms), C.byref(gemm_n), C.byref(all_n))
    L.gpb_profile_read_ozaki(C.byref(oz_ms), C.byref(oz_n), C.byref(oz_ops))
    L.gpb_profile_reset(0)
    value = n_total * steps / t
    # Flop per point (per rank share).  The public default (statistics="auto") takes the raw-product route while Kzz is well
    # conditioned: forward N M^2 (SYRK on [K_b|d|1]; the whitening is applied once to the M x M sums -- SURVEY 8d's
    # "accumulate Kzx Kxz first" variant, counted at its own, smaller figure) + backward 2 N M^2 (dK_b = [K_b|d|1] Caug^T).
    # The reference formulation (TRSM + SYRK forward, statistics="whitened") is 4 M^2.
    cond_est = sgpr_ops.kzz_condition_estimate(0, Z.detach(), ell.detach(), var.detach(), HYPER["jitter"])
    slot = sgpr_ops.route_state(0, Z, HYPER["jitter"])
    raw_route = bool(slot.decision) if slot is not None else False  # the route the timed steps actually took
    fpp = (3.0 if raw_route else 4.0) * m * m
    flops = fpp * (hi - lo)
    elbo = float(last["v"].item())
    phases = sgpr_ops.profile_phases(0, X, y, Z, ell, var, sn, c, HYPER["jitter"], block, None, raw_route)
    phases = {k: (D.max_over_ranks(v) if k != "elbo" else v) for k, v in phases.items()}
    int8_route = oz_n.value > 0
    # with the int8 route both streamed products (statistics SYRK, pass-2 dK_b) leave the DMMA pipe; what remains on it are the
    # replicated M x M finishes and blocks below the row threshold
    if int8_route:
        i8, bf16x2, peak_src = int8_roofline_peak(D)
        peak8 = i8["burst_tops"]
        a8 = oz_ops.value / (oz_ms.value * 1e-3) / 1e12
        roof = {"bound": "tensor",
                "kernel": "ozaki_i8_kernel (tcgen05.mma kind::i8, 7 radix-256 digit planes): K_b^T K_b statistics + dK_b = [K_b|d|1] Caug^T",
                "achieved": a8, "peak": peak8, "unit": "TFLOP/s", "frac": a8 / peak8, "peak_source": peak_src,
                "int8_ceiling_measured": i8, "frac_of_int8_sustained": a8 / i8["sustained_tops"], "frac_of_2x_bf16_sustained": a8 / bf16x2,
                "frac_of_nominal_dense_8bit_peak": a8 / NOMINAL_8BIT_TOPS,
                "int8_ops_per_point": oz_ops.value / steps / (hi - lo), "time_over_step_time": oz_ms.value * 1e-3 / t,
                "launches_per_step": oz_n.value / steps,
                "algorithmic_flop_per_point": fpp, "reference_formulation_flop_per_point": 4.0 * m * m,
                "statistics_route": "raw" if raw_route else "whitened", "kzz_condition_estimate": cond_est,
                "whole_step_tflops_per_gpu": flops * steps / t / 1e12,
                "whole_step_vs_fp64_dmma_peak": flops * steps / t / 1e12 / NOMINAL_FP64_TFLOPS,
                "remaining_dmma_gemms": {"time_over_step_time": gemm_ms.value * 1e-3 / t, "launches_per_step": gemm_n.value / steps},
                "traffic": None}
    else:
        achieved = flops * steps / (gemm_ms.value * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_f64_kernel (FP64 DMMA.8x8x4)", "achieved": achieved,
                "peak": NOMINAL_FP64_TFLOPS, "unit": "TFLOP/s", "frac": achieved / NOMINAL_FP64_TFLOPS,
                "algorithmic_flop_per_point": fpp, "reference_formulation_flop_per_point": 4.0 * m * m,
                "statistics_route": "raw" if raw_route else "whitened", "kzz_condition_estimate": cond_est,
                "gemm_time_over_step_time": gemm_ms.value * 1e-3 / t,
                "whole_step_tflops_per_gpu": flops * steps / t / 1e12, "traffic": None}

    # e2e through the public API with host-resident shards
    prior = gpx.gps.Prior(mean_function=gpx.mean_functions.Constant(Real(HYPER["mean_const"])),
                          kernel=gpx.kernels.RBF(lengthscale=PositiveReal(ell_ard(d)),
                                                 variance=NonNegativeReal(HYPER["variance"])))
    post = prior * gpx.likelihoods.Gaussian(num_datapoints=n_total, obs_stddev=NonNegativeReal(HYPER["obs_stddev"]))
    q = gpx.variational_families.CollapsedVariationalGaussian(posterior=post, inducing_inputs=Real(Zn),
                                                              jitter=HYPER["jitter"])
    leaves = [p for _, p in q.named_parameters()]
    host_out = {}

    def e2e_step():
        data = gpx.Dataset(X=Xh.to(D.dev, non_blocking=True), y=yh.to(D.dev, non_blocking=True))
        vals = [p.value.detach().requires_grad_(True) for p in leaves]
        for p, v in zip(leaves, vals):
            p.value = v
        loss = -gpx.objectives.collapsed_elbo(q, data, block_rows=block)
        grads = torch.autograd.grad(loss, vals)
        host_out["loss"] = loss.item()
        host_out["grads"] = [g.cpu() for g in grads]

    e2e_steps = 2
    t_e2e = timed(D, e2e_step, 1, e2e_steps)
    e2e = {"value": n_total * e2e_steps / t_e2e, "unit": "points/s", "h2d_bytes_per_step": int((hi - lo) * (d + 1) * 8),
           "d2h_bytes_per_step": int(8 * (1 + m * d + d + 3)), "steps": e2e_steps,
           "api": "gpx.objectives.collapsed_elbo(q, Dataset) + autograd, pinned host shard copied every step"}
    sgpr_ops.release_buffers()
    torch.cuda.empty_cache()
    return dict(metric="SGPR collapsed_elbo value+grad points/s", value=value, unit="points/s", n_gpus=D.world,
                steps=steps, warmup=warmup, ms_per_step=1e3 * t / steps, scaling="strong",
                config={"workload": f"sgpr_collapsed_elbo_value_and_grad_N{n_total}_M{m}_D{d}_RBF_ARD",
                        "N": n_total, "M": m, "D": d, "block_rows": block, "rows_per_rank": hi - lo,
                        "data": "one global synthetic set (65,536-row chunks seeded (4, chunk)); identical rows at every world size",
                        "collective": ("all-reduce (M+2)^2 fp64 fwd + (M*D+D+1) fp64 bwd: "
                                       + ("gpb_allreduce_f64 (NCCL through the C ABI, launching stream)" if native
                                          else "torch.distributed NCCL" if D.world > 1 else "none (1 rank)"))},
                elbo=elbo, elbo_note="same data at every world size: the values at 1/2/4/8 GPUs must agree to ~1e-11 relative",
                phases_ms_max_over_ranks={k: v for k, v in phases.items() if k != "elbo"}, elbo_from_phase_run=phases["elbo"],
                roofline=roof, e2e=e2e, gpu_launches=int(all_n.value))


def bench_svgp(D: Dist, args):
    """BASELINE config 5: SVGP minibatch ELBO value+grad, Matern32, N=50M (num_datapoints), D=16, M=4096,
    batch 65,536 per GPU drawn with replacement from the rank's shard (gpjax/fit.py:364-381), data-parallel."""
    torch = D.torch
    from gpjax_b200 import sgpr_ops, svgp_ops
    from gpjax_b200._lib import lib

    n_total, m, d = env_int("GPB_BENCH_SVGP_N", 50_000_000), env_int("GPB_BENCH_SVGP_M", 4096), 16
    batch = env_int("GPB_BENCH_SVGP_BATCH", 65536)
    shard = min(n_total // D.world, env_int("GPB_BENCH_SVGP_SHARD_ROWS", 4_000_000))  # resident part of the shard
    gen = torch.Generator(device=D.dev).manual_seed(5 + D.rank)
    X = torch.rand((shard, d), dtype=torch.float64, device=D.dev, generator=gen) * 4.0 - 2.0
    y = torch.sin(X[:, :1]) + 0.1 * torch.randn((shard, 1), dtype=torch.float64, device=D.dev, generator=gen)
    mk = lambda v: torch.as_tensor(np.asarray(v, np.float64), device=D.dev).requires_grad_(True)
    Z = mk(synth(m, d, 6)[0])
    ell, var, sn, c = mk(ell_ard(d)), mk(HYPER["variance"]), mk(HYPER["obs_stddev"]), mk(HYPER["mean_const"])
    mu, W = mk(np.zeros((m, 1))), mk(np.eye(m))
    params = (Z, ell, var, sn, c, mu, W)

    def step():
        for p in params:
            p.grad = None
        idx = torch.randint(0, shard, (batch,), device=D.dev, generator=gen)
        v = svgp_ops.svgp_elbo_fused(1, X[idx], y[idx], Z, ell, var, sn, c, mu, W, float(n_total), HYPER["jitter"], batch)
        v.backward()

    steps = max(2, args.steps)
    L = lib()
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    L.gpb_profile_reset(1)
    t = timed(D, step, 0, steps)
    import ctypes as C

    gemm_ms, gemm_n, all_n = C.c_double(), C.c_int64(), C.c_int64()
    oz_ms, oz_n, oz_ops = C.c_double(), C.c_int64(), C.c_double()
    L.gpb_profile_read(C.byref(gemm_ms), C.byref(gemm_n), C.byref(all_n))
    L.gpb_profile_read_ozaki(C.byref(oz_ms), C.byref(oz_n), C.byref(oz_ops))
    L.gpb_profile_reset(0)
    value = D.world * batch * steps / t
    # streamed passes (3 B M^2 on the raw-statistics route "auto" takes for a well-conditioned Kzz, else 4 B M^2)
    # + replicated M x M finish
    cond_est = sgpr_ops.kzz_condition_estimate(1, Z.detach(), ell.detach(), var.detach(), HYPER["jitter"])
    raw_route = cond_est <= sgpr_ops.RAW_STATISTICS_COND_LIMIT
    flops = (3.0 if raw_route else 4.0) * batch * m * m + 22.0 * m**3
    if oz_n.value > 0:  # streamed products on the int8 pipe (7 radix-256 digit planes); the M x M finish stays on DMMA
        i8, bf16x2, peak_src = int8_roofline_peak(D)
        peak8 = i8["burst_tops"]
        a8 = oz_ops.value / (oz_ms.value * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "ozaki_i8_kernel (tcgen05.mma kind::i8, 7 radix-256 digit planes)", "achieved": a8, "peak": peak8,
                "unit": "TFLOP/s", "frac": a8 / peak8, "peak_source": peak_src,
                "int8_ceiling_measured": i8, "frac_of_int8_sustained": a8 / i8["sustained_tops"], "frac_of_2x_bf16_sustained": a8 / bf16x2,
                "frac_of_nominal_dense_8bit_peak": a8 / NOMINAL_8BIT_TOPS,
                "time_over_step_time": oz_ms.value * 1e-3 / t, "algorithmic_flop_per_step": flops,
                "whole_step_tflops_per_gpu": flops * steps / t / 1e12,
                "remaining_dmma_gemms": {"time_over_step_time": gemm_ms.value * 1e-3 / t},
                "statistics_route": "raw" if raw_route else "whitened", "kzz_condition_estimate": cond_est, "traffic": None}
    else:
        roof = {"bound": "tensor", "kernel": "gemm_f64_kernel (FP64 DMMA.8x8x4)",
                "achieved": flops * steps / (gemm_ms.value * 1e-3) / 1e12, "peak": NOMINAL_FP64_TFLOPS,
                "unit": "TFLOP/s", "frac": flops * steps / (gemm_ms.value * 1e-3) / 1e12 / NOMINAL_FP64_TFLOPS,
                "algorithmic_flop_per_step": flops, "gemm_time_over_step_time": gemm_ms.value * 1e-3 / t,
                "statistics_route": "raw" if raw_route else "whitened", "kzz_condition_estimate": cond_est,
                "traffic": None}
    sgpr_ops.release_buffers()
    return dict(metric="SVGP elbo value+grad minibatch points/s", value=value, unit="points/s", n_gpus=D.world, steps=steps,
                ms_per_step=1e3 * t / steps, scaling="weak", higher_is_better=True, vs_baseline=None, dtype="f64",
                data="synthetic",
                config={"workload": f"svgp_elbo_value_and_grad_N{n_total}_M{m}_D{d}_B{batch}_Matern32", "N": n_total,
                        "M": m, "D": d, "batch_per_gpu": batch, "resident_shard_rows": shard},
                roofline=roof,
                gpu_launches=int(all_n.value))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "exact", "sgpr", "svgp"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    D = Dist(args.gpus)
    line = {}
    if args.workload in ("auto", "exact"):
        ex = bench_exact(D, args)
        line = {"metric": METRIC, "value": ex["value"], "unit": "evals/s", "n_gpus": D.world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * ex["t"] / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": ex["workload"], "N": ex["n"], "D": ex["d"], "kernel": "RBF ARD",
                           "parallelism": "replicas only (exact GP does not shard)" if D.world > 1 else "single GPU",
                           "l2": f"{8e-9 * ex['n'] ** 2:.1f} GB working set per step >> 126 MB L2 (no flush needed)",
                           "trailing_updates": ("int8 digit planes (Ozaki), default" if ex["roofline"].get("digit_planes")
                                                else "FP64 DMMA (GPB_OZAKI=0)")},
                "roofline": ex["roofline"], "clocks": ex["clocks"], "e2e": ex["e2e"], "gpu_launches": ex["gpu_launches"]}
    if args.workload in ("auto", "sgpr"):
        sg = bench_sgpr(D, args)
        if args.workload == "sgpr":
            line = {**sg, "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        else:
            line["sgpr"] = sg
    if args.workload == "svgp":
        line = bench_svgp(D, args)
    if D.rank == 0 and not args.no_cpu_baseline and D.world == 1:
        if args.workload in ("auto", "exact"):
            s = cpu_exact_sample(line["config"]["N"], 8)
            line["cpu_baseline"] = {"value": s["evals_per_s_at_target"], "unit": "evals/s", "cores": cpu_threads(),
                                    "kind": "port",
                                    "sample": f"one MLL+grad in the reference's LU formulation at N={s['n_sample']} "
                                              f"({s['seconds']:.2f} s), N^3-extrapolated to N={line['config']['N']}"}
        if args.workload in ("auto", "sgpr"):
            s = cpu_sgpr_sample(env_int("GPB_BENCH_SGPR_M", 2048), 8)
            cb = {"value": s["points_per_s"], "unit": "points/s", "cores": cpu_threads(), "kind": "port",
                  "sample": f"collapsed_elbo value+grad on {s['rows']} rows, M=2048 ({s['seconds']:.2f} s), linear in N"}
            if args.workload == "sgpr":
                line["cpu_baseline"] = cb
            else:
                line["sgpr"]["cpu_baseline"] = cb
    if D.rank == 0:
        print(json.dumps(line), flush=True)
    if D.dist is not None:
        D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
