#!/usr/bin/env python
"""Benchmark of the MU hot path: MU iterations/second on BASELINE.json config 3
(N=1024, T=2^20, K=32, L=64), T-sharded over --gpus GPUs of one box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--precision tf32|fp32] [--config C|B|D|E|A] [--t-scale S]

One "step" = one MultUpdate.update() (W terms, W update, reconstruction,
H terms, H update, reconstruction + loss).  Prints ONE JSON line (rank 0).
See DESIGN.md "Measurement" for how every field is derived.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests.cases import FULL, make_inputs  # noqa: E402

METRIC = "mu_iterations_per_second"
UNIT = "it/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CMF_BENCH_PRECISION", "auto"),
                    choices=["auto", "tf32", "tf32x3", "fp32"])
    ap.add_argument("--denominators", default=os.environ.get("CMF_BENCH_DENOMINATORS", "auto"),
                    choices=["auto", "direct", "gram"])
    ap.add_argument("--config", default="C", choices=sorted(FULL))
    ap.add_argument("--t-scale", type=float, default=1.0,
                    help="shrink T (debug only; the line is then labelled reduced)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile", action="store_true",
                    help="no per-kernel events in the timed region (lets the iteration replay as a CUDA graph)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-grade", action="store_true",
                    help="skip the short secondary measurement of the error-compensated tf32x3 mode")
    return ap.parse_args()


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Run this process on the CPUs local to GPU `index` (and so first-touch its pinned host buffers on that
    NUMA node): host<->device copies from the far socket run at a fraction of the PCIe rate.  Returns a short
    description for the JSON line, or None when the topology cannot be read."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not out:
            return None
        bdf = out[-12:] if len(out) >= 12 else out            # 0000:xx:yy.z (nvidia-smi prints an 8-digit domain)
        base = "/sys/bus/pci/devices/%s/" % bdf
        with open(base + "local_cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return "all %d allowed cpus are local" % len(allowed)
        os.sched_setaffinity(0, cpus)
        node = open(base + "numa_node").read().strip()
        return "numa node %s, %d cpus" % (node, len(cpus))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []
        self.t_begin = self.t_end = None

    def start(self):
        """Launch nvidia-smi and wait until it is actually sampling."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (ts, ln) in self.lines
                  if self.t_begin is None or (self.t_begin <= ts <= (self.t_end or ts) + 0.06)]
        if not inside:                        # region shorter than one sampling period
            inside = [ln for (_, ln) in self.lines[-3:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def cublas_tf32_tflops(torch, dev, n=8192, iters=10):
    """cuBLAS TF32 GEMM rate measured here and now (context for the roofline:
    MEASURED_PEAKS.json holds bf16 only).  Best of `iters`, CUDA events."""
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn((n, n), device=dev, dtype=torch.float32)
        b = torch.randn((n, n), device=dev, dtype=torch.float32)
        torch.matmul(a, b)
        best = 1e30
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


def algorithmic_flops(N, T, K, L):
    """six shift-contractions of 2 N K L T flops per iteration."""
    return 12.0 * N * K * L * T


def reference_iteration_seconds(N, T, K, L, n_iter, seed=0):
    """Times the oracle port of the reference's MU update (float64, per-lag
    NumPy/BLAS GEMMs, three reconstructions per iteration as the reference
    does) on this box's host cores."""
    from oracle import cmf_oracle
    X, W0, H0 = make_inputs(N, T, K, L, "uniform", seed)
    alg = cmf_oracle.MultUpdateOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64),
                                      initH=H0.astype(np.float64), tol=0, reuse_est=False)
    alg.update()                                   # warm-up (BLAS threads, page faults)
    t0 = time.perf_counter()
    for _ in range(n_iter):
        alg.update()
    return (time.perf_counter() - t0) / n_iter


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 0) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def cpu_baseline(N, T, K, L, budget_iters=2):
    """Bounded sample: the reference update at T_s = min(T, 16384) columns with
    identical N, K, L, extrapolated linearly in T (the reference's cost is
    linear in N*T*L)."""
    Ts = int(min(T, 16384))
    sec = reference_iteration_seconds(N, Ts, K, L, budget_iters)
    sec_full = sec * (T / Ts)
    return {"value": 1.0 / sec_full, "unit": UNIT, "cores": int(blas_threads()), "kind": "port",
            "sample": "oracle port of reference MultUpdate.update (float64 NumPy/BLAS), %d iterations at "
                      "N=%d T=%d K=%d L=%d (%.3f s/it), extrapolated linearly in T to T=%d"
                      % (budget_iters, N, Ts, K, L, sec, T)}


# --------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, T, K, L = FULL[args.config]
    T = int(T * args.t_scale)
    total = max(1, args.steps)
    from oracle import cmf_oracle
    # Sample size: as many columns as fit a ~2 minute run of `total` timed steps (longer samples are more faithful -
    # the per-column cost of the NumPy path grows once X leaves the CPU caches), between 4096 and 16384 columns.
    Ts = int(min(T, 4096))
    X, W0, H0 = make_inputs(N, Ts, K, L, "uniform", 0)
    probe = cmf_oracle.MultUpdateOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64),
                                        initH=H0.astype(np.float64), tol=0, reuse_est=False)
    t0 = time.perf_counter()
    probe.update()
    per_col = (time.perf_counter() - t0) / Ts
    while Ts * 2 <= min(T, 16384) and 1.4 * per_col * (Ts * 2) * (total + 2) <= 120.0:
        Ts *= 2
    del probe
    # warm-up + timed steps, each a bounded sample of the workload
    X, W0, H0 = make_inputs(N, Ts, K, L, "uniform", 0)
    alg = cmf_oracle.MultUpdateOracle(X.astype(np.float64), L, K, initW=W0.astype(np.float64),
                                      initH=H0.astype(np.float64), tol=0, reuse_est=False)
    for _ in range(max(1, min(args.warmup, 2))):
        alg.update()
    t0 = time.perf_counter()
    for _ in range(total):
        alg.update()
    sec = (time.perf_counter() - t0) / total
    sec_full = sec * (T / Ts)
    value = 1.0 / sec_full
    cores = int(blas_threads())
    sample = ("oracle port of reference MultUpdate.update (float64 NumPy/BLAS, %d BLAS threads; same per-lag "
              "GEMMs and three reconstructions as the reference but without its zero-pad copies, so it is "
              "faster than the unmodified reference), %d timed iterations at N=%d T=%d K=%d L=%d (%.3f s/it), "
              "extrapolated linearly in T to T=%d" % (cores, total, N, Ts, K, L, sec, T))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_full * 1e3,
            "higher_is_better": True, "scaling": "weak" if False else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config %s: N=%d T=%d K=%d L=%d MU" % (args.config, N, T, K, L)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def run_b200(args):
    import torch
    import __graft_entry__ as g
    g.build()
    from cmfpy_b200 import _lib
    lib = _lib.load()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    N, T, K, L = FULL[args.config]
    T = int(T * args.t_scale)
    assert T % world == 0
    Tloc = T // world
    precision = args.precision
    if precision == "auto":
        precision = "tf32" if lib.cmf_precision_supported(_lib.CMF_PREC_TF32, N, K, L) else "fp32"

    from cmfpy_b200.dist import ShardedMultUpdate
    # synthetic inputs generated on the device, identical for any world size:
    # global column t of X / H0 depends only on (seed, t)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    W0 = torch.rand((L, N, K), generator=gen, device=dev, dtype=torch.float32)
    chunk = 1 << 14
    X = torch.empty((N, Tloc + L - 1), device=dev, dtype=torch.float32)
    H0 = torch.empty((K, Tloc), device=dev, dtype=torch.float32)
    t_begin = rank * Tloc
    ncols_x = min(Tloc + L - 1, T - t_begin)
    for c0 in range(0, T, chunk):
        # every rank walks the same global stream so that shards agree
        xb = torch.rand((N, chunk), generator=gen, device=dev, dtype=torch.float32)
        hb = torch.rand((K, chunk), generator=gen, device=dev, dtype=torch.float32)
        lo, hi = max(c0, t_begin), min(c0 + chunk, t_begin + ncols_x)
        if lo < hi:
            X[:, lo - t_begin:hi - t_begin] = xb[:, lo - c0:hi - c0]
        lo, hi = max(c0, t_begin), min(c0 + chunk, t_begin + Tloc)
        if lo < hi:
            H0[:, lo - t_begin:hi - t_begin] = hb[:, lo - c0:hi - c0]
    del xb, hb
    # scale the init like rand_init does, cheaply: E[X]=0.5, E[est]=L*K/4
    s = float(np.sqrt(0.5 / (L * K / 4.0)))
    W0 *= s
    H0 *= s

    alg = ShardedMultUpdate(X[:, :ncols_x], N, T, K, L, t_offset=t_begin, t_local=Tloc,
                            initW=W0, initH=H0, precision=precision, device=local_rank,
                            group=dist.group.WORLD if dist else None, denominators=args.denominators)
    torch.cuda.synchronize()

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    alg.update_many(args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = alg.launch_count
    alg.set_profiling(not args.no_profile)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = alg.torch_stream
    barrier()
    sampler.mark_begin()
    ev0.record(stream)
    losses = alg.update_many(args.steps)
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    kms = alg.kernel_ms()
    alg.set_profiling(False)
    launches = alg.launch_count - launches0
    if dist:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = args.steps / (ms * 1e-3)

    # ---- end-to-end through the public host API (pinned host buffers) -------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, N, T, K, L, Tloc, rank, world, local_rank, precision, dist)

    # fp32-grade companion number (north_star: the 1e-4 parity bar is an fp32 bar, "TF32 variant reported
    # separately"): the same workload on the error-compensated tensor-core mode, a few steps, outside the timed region
    fp32_grade = None
    if (precision == "tf32" and not args.no_fp32_grade and
            lib.cmf_precision_supported(_lib.CMF_PREC_TF32X3, N, K, L)):
        alg3 = ShardedMultUpdate(X[:, :ncols_x], N, T, K, L, t_offset=t_begin, t_local=Tloc,
                                 initW=W0, initH=H0, precision="tf32x3", device=local_rank,
                                 group=dist.group.WORLD if dist else None)
        alg3.update_many(2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(alg3.torch_stream)
        l3 = alg3.update_many(5)
        e1.record(alg3.torch_stream)
        torch.cuda.synchronize()
        ms3 = e0.elapsed_time(e1)
        if dist:
            t3 = torch.tensor([ms3], device=dev)
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            ms3 = float(t3.item())
        fp32_grade = {"precision": "tf32x3", "path": alg3.path_name, "value": 5 / (ms3 * 1e-3), "unit": UNIT,
                      "steps": 5, "warmup": 2, "loss_after_7_steps": l3[-1],
                      "what": "same workload, every product as three TF32 MMAs on hi/lo operand pairs; loss "
                              "trajectories within 1e-4 of the float64 reference on all golden cases but config B "
                              "(profiles/r01_trajectory_errors_tf32x3.log)"}
        alg3.close()
        del alg3
    del X, H0
    if rank != 0:
        alg.close()
        if dist:
            dist.destroy_process_group()
        return

    peaks, peaks_src = measured_peaks()
    tf32_live = cublas_tf32_tflops(torch, dev)
    flops_iter = algorithmic_flops(N, T, K, L)
    # reconstructions per step: two with the direct denominators, one with the Gram route (est is then
    # needed for the loss only)
    recon_launches = (1 if alg.path_name.endswith("+gram") else 2) * args.steps
    recon_ms = max(kms["recon"] / recon_launches, 1e-9)
    recon_flops = 2.0 * N * K * L * Tloc            # per launch, per GPU
    achieved = recon_flops / (recon_ms * 1e-3) / 1e12 if kms["recon"] > 0 else None
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    roofline = {
        "bound": "tensor", "kernel": "recon (shift-GEMM, %s)" % alg.path_name,
        "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s",
        "frac": (achieved / tf32_peak) if achieved else None,
        # DRAM bytes per K1 launch from `ncu --set full` (profiles/r01_ncu_tc_kernels_gram.txt /
        # _direct.txt); captured for config C on one GPU only
        "traffic": ((4.44e9 if alg.path_name.endswith("+gram") else 8.70e9)
                    if (args.config == "C" and world == 1 and args.t_scale == 1.0 and precision == "tf32") else None),
        "traffic_unit": "bytes per launch (algorithmic: %.2e)" % ((4.0 if alg.path_name.endswith("+gram") else 8.0) * N * Tloc),
        "peak_source": "%s bf16_tflops_sustained / 2 (TF32 dense = half of bf16; not separately measured)" % peaks_src,
        "cublas_tf32_tflops_live": tf32_live,
        # the tensor pipe itself: one 128x256x8 TF32 MMA (524288 flop) per 128 cycles per SM at the SM clock
        # observed DURING the timed region (the run is power-capped well below the 1965 MHz maximum)
        "hw_tf32_tflops_at_observed_clock": (148 * 4096.0 * clocks["sm_mhz"] * 1e6 / 1e12
                                             if clocks and clocks.get("sm_mhz") else None),
        "frac_of_hw_at_observed_clock": (achieved / (148 * 4096.0 * clocks["sm_mhz"] * 1e6 / 1e12)
                                         if achieved and clocks and clocks.get("sm_mhz") else None),
        # 12 N K L T per iteration (what the direct algorithm needs) over the measured time; with
        # the Gram-route denominators about half of those flops are not executed at all, so this figure is a
        # reference-equivalent rate, not a hardware utilisation
        "reference_equivalent_tflops": flops_iter / world / (ms_per_step * 1e-3) / 1e12,
        "executed_tflops_approx": (flops_iter * (0.5 + (2.0 * K + 4.0 * K * (2 * L - 1) / L) / (6.0 * N))
                                   if alg.path_name.endswith("+gram") else
                                   (3.0 * flops_iter if precision == "tf32x3" else flops_iter))
                                  / world / (ms_per_step * 1e-3) / 1e12,
        "kernel_ms_per_step": {k: v / args.steps for k, v in kms.items()},
        "hbm_update_kernels": {
            # W and H multiplicative updates: 16 B/element algorithmic (+4 B for the TF32 operand copy)
            "bytes_per_step": (20 if precision == "tf32" else 16) * (L * N * K + K * Tloc),
            "achieved_gbs": (20 if precision == "tf32" else 16) * (L * N * K + K * Tloc) /
                            max(kms["elementwise"] / args.steps * 1e-3, 1e-12) / 1e9,
            "peak_gbs": peaks["hbm_gbs"]},
    }
    if world > 1:
        # the "elementwise" interval of a sharded step also holds the exchange kernels and their waits for the
        # slowest rank: the update kernels are not isolated there, so no bandwidth is claimed for them
        roofline["hbm_update_kernels"] = None
    cb = None
    if not args.no_cpu_baseline and world == 1:          # the CPU baseline is a 1-GPU-run item (rank 0, N = 1 only)
        cb = cpu_baseline(N, T, K, L)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": {"tf32": "tf32", "tf32x3": "tf32x3 (3 TF32 MMAs per product, fp32-grade)"}.get(precision, "f32"), "data": "synthetic",
        "config": {"workload": "config %s: N=%d T=%d K=%d L=%d MU%s" %
                   (args.config, N, T, K, L, "" if args.t_scale == 1.0 else " (T reduced: debug)"),
                   "sharding": "time axis, %d x %d columns, halo %d" % (world, Tloc, L - 1),
                   "l2": "inputs_exceed_l2 (X and est are %.1f GiB each per GPU)" % (N * Tloc * 4 / 2**30),
                   "precision": precision, "denominators": args.denominators, "path": alg.path_name,
                   "collectives": ("none (1 GPU)" if world == 1 else
                                   {"peer": "own kernels over NVLink peer memory (all-reduce fused with the W update, "
                                            "halo pushes, loss ring)",
                                    "nccl": "NCCL all-reduce + send/recv between the phases"}[alg.transport])},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "cpu_affinity": affinity,
        "roofline": roofline, "cpu_baseline": cb, "fp32_grade": fp32_grade,
        "final_loss": losses[-1],
    }
    print(json.dumps(line), flush=True)
    alg.close()
    if dist:
        dist.destroy_process_group()


def run_e2e(args, torch, N, T, K, L, Tloc, rank, world, local_rank, precision, dist):
    """The call a user makes: solver built from HOST arrays (pinned), every
    update() returns its loss to the host, W and H read back at the end.  The
    timed region holds all host<->device traffic."""
    from cmfpy_b200.dist import ShardedMultUpdate
    dev = torch.device("cuda", local_rank)
    t_begin = rank * Tloc
    ncols_x = min(Tloc + L - 1, T - t_begin)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99 + rank)
    Xh = torch.empty((N, ncols_x), dtype=torch.float32, pin_memory=True)
    Xh.copy_(torch.rand((N, ncols_x), generator=gen, device=dev))
    s = float(np.sqrt(0.5 / (L * K / 4.0)))
    gen.manual_seed(7)
    W0h = torch.empty((L, N, K), dtype=torch.float32, pin_memory=True)
    W0h.copy_(torch.rand((L, N, K), generator=gen, device=dev) * s)
    gen.manual_seed(99 + rank)
    H0h = torch.empty((K, Tloc), dtype=torch.float32, pin_memory=True)
    H0h.copy_(torch.rand((K, Tloc), generator=gen, device=dev) * s)
    # results land in pinned host buffers too (a direct DMA; a fresh pageable array costs a staged copy plus the
    # page faults of 128 MiB of new memory - 0.04 to 0.8 s on these VMs)
    Wout = torch.empty((L, N, K), dtype=torch.float32, pin_memory=True)
    Hout = torch.empty((K, Tloc), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    alg = ShardedMultUpdate(Xh.numpy(), N, T, K, L, t_offset=t_begin, t_local=Tloc,
                            initW=W0h.numpy(), initH=H0h.numpy(), precision=precision,
                            device=local_rank, group=dist.group.WORLD if dist else None,
                            denominators=args.denominators)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = alg.update()                       # host float every step (D2H + sync)
    t2 = time.perf_counter()
    W = alg.W_host(out=Wout.numpy())
    H = alg.H_local_host(out=Hout.numpy())
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sec = time.perf_counter() - t0
    parts = {"construct_h2d": t1 - t0, "steps": t2 - t1, "read_back": time.perf_counter() - t2}
    if dist:
        t = torch.tensor([sec], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    h2d = (Xh.numel() + W0h.numel() + H0h.numel()) * 4
    d2h = (W.size + H.size) * 4 + 8 * args.steps
    alg.close()
    return {"value": args.steps / sec, "unit": UNIT,
            "h2d_bytes_per_step": int(h2d * world / args.steps),
            "d2h_bytes_per_step": int(d2h * world / args.steps),
            "seconds_total": sec, "seconds_rank0": parts, "final_loss": last,
            "what": "solver built from pinned host X/W0/H0 (H2D inside the timed region), %d update() calls each "
                    "returning the loss to the host, W and H copied back into pinned host arrays" % args.steps}


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries underneath (NCCL's
    version banner, for one) write to file descriptor 1 directly, so fd 1 is
    pointed at stderr for the whole run and the JSON line goes to the saved
    original descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(saved, "w")
    real_print = print

    def emit(*args, **kw):
        kw.pop("flush", None)
        real_print(*args, file=out, **kw)
        out.flush()
    return emit


if __name__ == "__main__":
    a = parse_args()
    emit = _claim_stdout()
    import builtins
    builtins.print = emit           # rank 0 prints exactly one line; nothing else reaches stdout
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
