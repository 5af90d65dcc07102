#!/usr/bin/env python
"""bench.py — KKT factor+solve throughput of the cone-program hot path on B200.

A "step" is one interior-point iteration's KKT work on the north-star workload
(dense QP n=8192, m=2n 'l' cone rows): 1 factor (NT scaling fused into the
normal-equations SYRK + Cholesky) and 2 solves (affine + combined direction;
coneqp with refinement 0, reference src/python/coneprog.py:2256, 2360-2401).

  value  : algorithmic GF/s (F_it / step time) with every input resident
           in HBM, timed with CUDA events on the library's launch stream.
  e2e    : the same metric through the public plugin call a CVXOPT user makes
           (cvxopt_b200.kkt_chol -> factor(W) -> solve(x,y,z)) with pinned HOST buffers;
           H2D/D2H copies inside the timed region.
  --impl reference : the unmodified reference path (oracle/_ref: cvxopt's misc.kkt_chol +
           OpenBLAS) on the box's host cores, same workload and metric.

N > 1 GPUs: a single factorisation does not shard -> N independent
replicas, one per rank, weak scaling; value = total flops / max-over-ranks time.
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

METRIC = "kkt_factor_solve_gflops_fp64"
INT8_TENSOR_PEAK_TOPS = 4500.0      # nominal dense int8 (B200); no measured entry yet
FP64_DMMA_PEAK_TFLOPS = 37.2   # tools/fp64_peak.cu on this pool's B200 (profiles/r01_fp64_peaks.md)


def flops(n, m, refinement=0):
    f_fac = float(n) * n * m + float(n) ** 3 / 3.0          # SYRK n^2 Kp + POTRF n^3/3
    f_sol = 4.0 * m * n + 2.0 * float(n) * n
    return f_fac, f_sol, f_fac + 2 * (1 + refinement) * f_sol


def make_problem(n, m, seed):
    """SURVEY.md §8(d) dense QP generator (numpy PCG64)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    A0 = rng.standard_normal((n, n))
    P = np.asfortranarray(A0.T @ A0 / n + np.eye(n))
    del A0
    G = np.asfortranarray(rng.standard_normal((n, m)).T)     # m x n, column-major
    # mid-IPM scaling: d spans ~4 decades
    d = 10.0 ** rng.uniform(-2.0, 2.0, m)
    return P, G, d, rng


def make_qp(n, m, seed):
    """full dense QP: P, q, G, h with a strictly feasible point"""
    rng = np.random.Generator(np.random.PCG64(seed))
    A0 = rng.standard_normal((n, n))
    P = A0.T @ A0 / n + np.eye(n)
    del A0
    q = rng.standard_normal(n)
    G = rng.standard_normal((m, n))
    x0 = rng.standard_normal(n)
    h = G @ x0 + rng.uniform(0.1, 1.1, m)
    return P, q, G, h


def run_ipm(n, m, seed, device):
    """whole interior-point solve of one dense QP, device resident (cvxopt_b200.QPBatch, B=1)"""
    import cvxopt_b200
    P, q, G, h = make_qp(n, m, seed)
    b = cvxopt_b200.QPBatch(1, n, m, device)
    b.load(P[None], q[None], G[None], h[None])
    b.solve()                      # warm-up (first-launch overheads, clocks)
    t0 = time.perf_counter()
    b.solve()
    wall = (time.perf_counter() - t0) * 1e3
    r, st = b.results(), b.stats()
    b.close()
    it = int(r["iterations"][0])
    f_fac, f_sol, f_it = flops(n, m)
    return {"n": n, "m": m, "iterations": it, "status": r["status"][0], "ms_total": st["solve_ms"],
            "wall_ms": wall, "iters_per_s": (it + 1) / (st["solve_ms"] * 1e-3),
            "primal_objective": float(r["primal objective"][0]),
            "gflops": (it + 1) * f_it / (st["solve_ms"] * 1e-3) * 1e-9}


def run_batch(nprob, n, m, device, rank, world):
    """BASELINE config 4: nprob independent QPs sharded over the ranks (contiguous blocks)"""
    import cvxopt_b200
    lo, hi = cvxopt_b200.shard_bounds(nprob, world)[rank]
    Ps, qs, Gs, hs = [], [], [], []
    for k in range(lo, hi):
        P, q, G, h = make_qp(n, m, k)
        Ps.append(P); qs.append(q); Gs.append(G); hs.append(h)
    b = cvxopt_b200.QPBatch(hi - lo, n, m, device)
    b.load(np.stack(Ps), np.stack(qs), np.stack(Gs), np.stack(hs))
    b.solve()
    b.solve()
    r, st = b.results(), b.stats()
    b.close()
    return {"problems": hi - lo, "ms": st["solve_ms"], "lockstep_iterations": st["lockstep_iterations"],
            "iterations_sum": int(r["iterations"].sum()),
            "all_optimal": bool(all(x == "optimal" for x in r["status"]))}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4)
                          if len(s) >= 7 and s[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def run_reference(args, rank, world):
    """CPU arm: the reference's own kkt_chol (oracle/_ref) on host cores, same step."""
    if rank != 0:
        return
    n, m = args.n, args.m
    f_fac, f_sol, f_it = flops(n, m)
    base = {"metric": METRIC, "unit": "GF/s", "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "dense QP KKT step n=%d m=%d: 1 factor + 2 solves (kkt_chol)" % (n, m)}}
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "cvxopt")):
        base["unavailable"] = "oracle/_ref not built (oracle/build_ref.sh needs /root/reference)"
        print(json.dumps(base))
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(min(cores, 64)))
    sys.path.insert(0, ref_dir)
    from cvxopt import matrix, misc
    P, G, d, rng = make_problem(n, m, args.seed)
    Gm, Pm = matrix(G), matrix(P)
    dims = {"l": m, "q": [], "s": []}
    factor = misc.kkt_chol(Gm, dims, matrix(0.0, (0, n)))
    W = {"d": matrix(d), "di": matrix(1.0 / d), "v": [], "beta": [], "r": [], "rti": []}
    y = matrix(0.0, (0, 1))

    def step():
        f = factor(W, Pm)
        for _ in range(2):
            x, z = matrix(rng.standard_normal(n)), matrix(rng.standard_normal(m))
            f(x, y, z)
    # a full-size reference step takes tens of seconds on the host (single-threaded glue in
    # misc.kkt_chol.factor): bound the sample so the whole arm ends within a few minutes
    est = 2.4e-11 * f_it                      # ~ seconds per step at the ~45 GF/s measured on this pool
    steps_run = max(1, min(args.steps, int(60.0 / max(est, 1e-3)) or 1))
    warm_run = max(1, min(args.warmup, int(30.0 / max(est, 1e-3)) or 1))
    for _ in range(warm_run):
        step()
    t0 = time.perf_counter()
    for _ in range(steps_run):
        step()
    dt = time.perf_counter() - t0
    ms = dt / steps_run * 1e3
    val = f_it / (ms * 1e-3) * 1e-9
    base.update({"value": val, "ms_per_step": ms,
                 "cpu_baseline": {"value": val, "unit": "GF/s", "cores": min(cores, 64), "kind": "reference",
                                  "sample": "%d full-size steps after %d warm-up (misc.kkt_chol, scipy-openblas)"
                                            % (steps_run, warm_run)},
                 "steps_run": steps_run,
                 "e2e": {"value": val, "unit": "GF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(base))


def cpu_baseline_sample(args, P, G, d):
    """Bounded sample of the reference path on this box's host cores (rank 0, N=1)."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "cvxopt")):
        return {"value": None, "unit": "GF/s", "cores": 0, "kind": "reference",
                "sample": "oracle/_ref not built"}
    code = r"""
import os, sys, json, time
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import bench
from cvxopt import matrix, misc
n, m, seed = %d, %d, %d
P, G, d, rng = bench.make_problem(n, m, seed)
factor = misc.kkt_chol(matrix(G), {'l': m, 'q': [], 's': []}, matrix(0.0, (0, n)))
W = {'d': matrix(d), 'di': matrix(1.0 / d), 'v': [], 'beta': [], 'r': [], 'rti': []}
Pm, y = matrix(P), matrix(0.0, (0, 1))
def step():
    f = factor(W, Pm)
    for _ in range(2):
        f(matrix(rng.standard_normal(n)), y, matrix(rng.standard_normal(m)))
# warm-up (OpenBLAS thread start-up, first-touch) on a small instance of the same code path: a
# full-size reference step takes 30-70 s on this pool's hosts
Pw, Gw, dw, rw = bench.make_problem(512, 1024, seed)
fw = misc.kkt_chol(matrix(Gw), {'l': 1024, 'q': [], 's': []}, matrix(0.0, (0, 512)))
fw({'d': matrix(dw), 'di': matrix(1.0 / dw), 'v': [], 'beta': [], 'r': [], 'rti': []}, matrix(Pw))(
    matrix(rw.standard_normal(512)), y, matrix(rw.standard_normal(1024)))
t0 = time.perf_counter(); k = 0
while k < 1 or (time.perf_counter() - t0 < 10.0 and k < 20):
    step(); k += 1
print(json.dumps({'ms': (time.perf_counter() - t0) / k * 1e3, 'steps': k}))
""" % (ref_dir, ROOT, args.n, args.m, args.seed)
    cores = os.cpu_count() or 1
    env = dict(os.environ, OPENBLAS_NUM_THREADS=str(min(cores, 64)))
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                             timeout=600)
        res = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        return {"value": None, "unit": "GF/s", "cores": min(cores, 64), "kind": "reference",
                "sample": "failed: %s" % e}
    _, _, f_it = flops(args.n, args.m)
    return {"value": f_it / (res["ms"] * 1e-3) * 1e-9, "unit": "GF/s", "cores": min(cores, 64),
            "kind": "reference", "ms_per_step": res["ms"],
            "sample": "%d full-size step(s) (n=%d, m=%d) after a small warm-up, reference misc.kkt_chol "
                      "on scipy-openblas, %d threads" % (res["steps"], args.n, args.m, min(cores, 64))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nvars", "--n", dest="n", type=int, default=8192)
    ap.add_argument("--mrows", "--m", dest="m", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ipm", action="store_true", help="skip the full-IPM and batch extras")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--no-i8", action="store_true", help="skip the extra leg on the experimental int8-slice SYRK")
    args = ap.parse_args()
    if args.m <= 0:
        args.m = 2 * args.n
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import cvxopt_b200
    from cvxopt_b200 import _lib
    if not torch.cuda.is_available() or cvxopt_b200.device_count() == 0:
        raise RuntimeError("bench.py needs a B200: no CUDA device visible (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, m = args.n, args.m
    # csrc/kkt_api.cu: the 'l'-row SYRK of large problems runs on the int8 tensor path unless CVXB_OZAKI=0
    oz_env = os.environ.get("CVXB_OZAKI", "1")[:1]
    i8_default = (oz_env == "2") or (oz_env == "1" and n >= 4096 and m >= 8192)
    f_fac, f_sol, f_it = flops(n, m)
    P, G, d, rng = make_problem(n, m, args.seed + rank)
    dims = {"l": m, "q": [], "s": []}
    kkt = cvxopt_b200.kkt_chol(G, dims, None, H=P, device=local_rank)

    # ---------------- device-resident arm (value) ----------------
    dev = torch.device("cuda", local_rank)
    d_d = torch.from_numpy(d).to(dev)
    d_di = torch.from_numpy(1.0 / d).to(dev)
    xs = [torch.from_numpy(rng.standard_normal(n)).to(dev) for _ in range(2)]
    zs = [torch.from_numpy(rng.standard_normal(m)).to(dev) for _ in range(2)]
    xw, zw = torch.empty_like(xs[0]), torch.empty_like(zs[0])

    def step_dev():
        kkt.factor_ptr(d=d_d.data_ptr(), di=d_di.data_ptr(), space=_lib.DEVICE)
        for i in range(2):
            xw.copy_(xs[i]); zw.copy_(zs[i])
            torch.cuda.current_stream().synchronize()
            kkt.solve_ptr(xw.data_ptr(), zw.data_ptr(), space=_lib.DEVICE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_dev()
    syrk_ms, potrf_ms, fac_ms, sol_ms = [], [], [], []
    barrier()
    launches0 = cvxopt_b200.launch_count()
    with ClockSampler(local_rank) as clk:
        kkt.timer_start()
        for _ in range(args.steps):
            step_dev()
            b = kkt.last_breakdown()
            syrk_ms.append(b["syrk_ms"]); potrf_ms.append(b["potrf_ms"])
            f, s = kkt.last_ms()
            fac_ms.append(f); sol_ms.append(s)
        total_ms = kkt.timer_stop()
    launches = cvxopt_b200.launch_count() - launches0
    barrier()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * f_it / (ms_step * 1e-3) * 1e-9

    # ---------------- end-to-end arm (host buffers through the plugin API) ----------------
    def pinned(a):
        tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return tt, tt.numpy()
    keep = []
    h_d = pinned(d); h_di = pinned(1.0 / d)
    keep += [h_d, h_di]
    W = {"d": h_d[1], "di": h_di[1], "v": [], "beta": [], "r": [], "rti": []}
    hx = [pinned(rng.standard_normal(n)) for _ in range(2)]
    hz = [pinned(rng.standard_normal(m)) for _ in range(2)]
    hxw, hzw = pinned(np.zeros(n)), pinned(np.zeros(m))

    def step_e2e():
        solve = kkt(W)                       # f = kktsolver(W)
        for i in range(2):
            hxw[1][:] = hx[i][1]; hzw[1][:] = hz[i][1]
            solve(hxw[1], None, hzw[1])      # f(x, y, z), in place on host buffers
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_val = world * f_it / (e2e_ms * 1e-3) * 1e-9
    h2d = 2 * m * 8 + 2 * (n + m) * 8
    d2h = 2 * (n + m) * 8

    if rank == 0:
        syrk = float(np.mean(syrk_ms))
        f_syrk = float(n) * n * m
        achieved = f_syrk / (syrk * 1e-3) * 1e-12
        if i8_default:
            # the 'l'-row SYRK ran as 45 exact int8 products (nine radix-2^7 slices per entry) on
            # tcgen05.mma kind::i8: the bounding pipe is the int8 tensor pipe.  syrk_ms covers the slicing
            # kernels (~1 ms) as well as oz_mma_kernel, so `achieved` is a lower bound for the kernel alone.
            tiles = ((n + 127) // 128) * ((n + 127) // 128 + 1) // 2
            i8_ops = 45.0 * 2.0 * tiles * 128.0 * 128.0 * m
            a8 = i8_ops / (syrk * 1e-3) * 1e-12
            roofline = {"kernel": "oz_mma_kernel (int8-slice SYRK: tcgen05.mma.kind::i8, int32 accumulators in TMEM)",
                        "bound": "tensor", "achieved": a8, "peak": INT8_TENSOR_PEAK_TOPS, "unit": "TFLOP/s",
                        "frac": a8 / INT8_TENSOR_PEAK_TOPS,
                        "peak_source": "nominal dense int8 rate of B200 (4.5 POP/s; MEASURED_PEAKS.json has no int8 entry); "
                                       "`achieved` counts int8 multiply-add ops of the 45 slice products",
                        "fp64_equivalent_tflops": achieved, "fp64_dmma_peak_tflops": FP64_DMMA_PEAK_TFLOPS,
                        # dram__bytes_read.sum + dram__bytes_write.sum of one oz_mma_kernel launch from the
                        # committed ncu capture (profiles/r01k_ozaki_syrk_ncu_summary.md), n=8192 only
                        "traffic": (17.30e9 + 0.81e9) if (n == 8192 and m == 16384) else None,
                        "algorithmic_bytes": 8.0 * m * n + 9.0 * m * n * 2 + 8.0 * n * n}
        else:
            roofline = {"kernel": "dmma_gemm_kernel<XK,YK,VEC> (fused NT-scaled SYRK)", "bound": "tensor",
                        "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                        "frac": achieved / FP64_DMMA_PEAK_TFLOPS,
                        "peak_source": "measured DMMA.8x8x4 pipe rate on this pool (tools/fp64_peak.cu; "
                                       "MEASURED_PEAKS.json has no fp64 entry; cuBLAS DGEMM 8192^3 = 35.4)",
                        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the
                        # committed capture profiles/r01j_syrk_band_order_dram.md (n=8192 only)
                        "traffic": (7.30e9 + 0.275e9) if (n == 8192 and m == 16384) else None,
                        "algorithmic_bytes": 8.0 * m * n + 8.0 * n * n}
        out = {
            "metric": METRIC, "value": value, "unit": "GF/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "dense QP KKT step n=%d m=%d ('l' cone): 1 factor + 2 solves" % (n, m),
                       "parallelism": "replicas x%d" % world if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (G = %.2f GB, K = %.2f GB)" % (8.0 * n * m / 1e9, 8.0 * n * n / 1e9),
                       "flops_per_step": f_it,
                       "syrk_path": ("9 int8 slices per entry on tcgen05.mma kind::i8 (exact int32 products, fp64 recombination)"
                                     if i8_default else "fp64 DMMA"),
                       "precision": ("fp64 results: the slice products are exact integers and 9 slices keep 62 bits below each "
                                     "column maximum; error vs an 80-bit evaluation 4e-16 * sum|terms|, the fp64 dot-product level "
                                     "(tests/test_i8_syrk_gpu.py, profiles/r01k); the all-fp64 path is timed as other_syrk_path")
                                    if i8_default else "fp64 throughout"},
            "e2e": {"value": e2e_val, "unit": "GF/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "breakdown_ms": {"factor": float(np.mean(fac_ms)), "syrk": syrk, "potrf": float(np.mean(potrf_ms)),
                             "solve_each": float(np.mean(sol_ms))},
            "ipm_iters_per_s_kkt_bound": 1e3 / ms_step,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample(args, P, G, d)
    extras = {}
    if rank == 0 and world == 1 and not args.no_i8:
        # extra leg (not the headline): the same KKT step with the 'l'-row SYRK on the other tensor path
        # (fp64 DMMA kernel when the int8-slice kernel is the default, and vice versa), timed the same way,
        # and its direction against the default path's
        saved = os.environ.get("CVXB_OZAKI")
        os.environ["CVXB_OZAKI"] = "0" if i8_default else "2"
        kkt8 = cvxopt_b200.kkt_chol(G, dims, None, H=P, device=local_rank)
        if saved is None:
            os.environ.pop("CVXB_OZAKI")
        else:
            os.environ["CVXB_OZAKI"] = saved

        def step8():
            kkt8.factor_ptr(d=d_d.data_ptr(), di=d_di.data_ptr(), space=_lib.DEVICE)
            for i in range(2):
                xw.copy_(xs[i]); zw.copy_(zs[i])
                torch.cuda.current_stream().synchronize()
                kkt8.solve_ptr(xw.data_ptr(), zw.data_ptr(), space=_lib.DEVICE)
        for _ in range(3):
            step8()
        x8 = xw.clone()
        step_dev()
        xdiff = float((torch.linalg.vector_norm(x8 - xw) / torch.linalg.vector_norm(xw)).item())
        torch.cuda.synchronize()
        k8 = max(3, min(args.steps, 10))
        s8 = []
        kkt8.timer_start()
        for _ in range(k8):
            step8()
            s8.append(kkt8.last_breakdown()["syrk_ms"])
        ms8 = kkt8.timer_stop() / k8
        extras["other_syrk_path"] = {
            "what": ("fp64 DMMA SYRK (CVXB_OZAKI=0)" if i8_default else "int8-slice SYRK on tcgen05.mma kind::i8 (CVXB_OZAKI=2)")
                    + ": same step, same inputs",
            "ms_per_step": ms8, "value": f_it / (ms8 * 1e-3) * 1e-9, "syrk_ms": float(np.mean(s8)),
            "syrk_fp64_equiv_tflops": float(n) * n * m / (float(np.mean(s8)) * 1e-3) * 1e-12,
            "direction_rel_diff_vs_default": xdiff}
        kkt8.close()
        del kkt8
    kkt.close()
    del kkt, d_d, d_di, xs, zs
    torch.cuda.empty_cache()
    if not args.no_ipm:
        # config 4: batch of independent QPs, strong scaling over the ranks
        bres = run_batch(args.batch, 512, 1024, local_rank, rank, world)
        t = torch.tensor([bres["ms"]], dtype=torch.float64, device=dev)
        cnt = torch.tensor([float(bres["iterations_sum"]), float(bres["problems"]), float(bres["all_optimal"])],
                           dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        if rank == 0:
            bf = flops(512, 1024)[2]
            extras["batch"] = {"workload": "%d independent dense QPs n=512 m=1024, %d per GPU" % (args.batch, bres["problems"]),
                               "ms": float(t.item()), "problems_per_s": args.batch / (float(t.item()) * 1e-3),
                               "ipm_iterations_total": int(cnt[0].item()),
                               "gflops": (cnt[0].item() + cnt[1].item()) * bf / (float(t.item()) * 1e-3) * 1e-9,
                               "all_optimal": bool(cnt[2].item() == world), "scaling": "strong"}
        # (ii) IPM iterations/s: a whole solve of the same-size QP, device resident (rank 0 only, last leg:
        # nothing after it depends on it).  The device IPM's factor uses the fp64 DMMA SYRK (its int8-slice
        # call site is opt-in, CVXB_OZAKI_IPM=1, until it has been validated on a GPU).
        if rank == 0:
            try:
                ipm = run_ipm(n, m, args.seed, local_rank)
                ipm["syrk_path"] = "int8 slices" if os.environ.get("CVXB_OZAKI_IPM", "0")[:1] in ("1", "2") else "fp64 DMMA"
                extras["ipm"] = ipm
            except Exception as exc:        # keep the headline line even if this extra leg fails
                extras["ipm"] = {"error": repr(exc)[:300]}
    if rank == 0:
        out.update(extras)
        print(json.dumps(out))
    if world > 1:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()
