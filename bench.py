"""Benchmark of the hot path: exact k-NN search + IW entropy forward/backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): k-NN entropy states/sec (search + IW fwd/bwd) at N=1M, d=29, k=4.
One "step" = one search of all N synthetic particles (k+1=5, self included), the reverse-CSR build
for that search, and one IW-entropy forward (H + KL) + backward on the same N states.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "k-NN entropy states/sec (search + IW fwd/bwd)"
UNIT = "states/s"
DIGAMMA = {1: -0.5772156649015329, 4: 1.2561176684318003, 50: 3.9019896734278921}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1048576)
    ap.add_argument("--d", type=int, default=29)
    ap.add_argument("--k", type=int, default=4)
    ap.add_argument("--traj-len", type=int, default=500)
    ap.add_argument("--algo", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline sample time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def synth(n, d, traj_len, seed=0):
    """Synthetic batch of BASELINE.md section 4: N(0,1) float32 particles (tie-free), trajectories of
    length traj_len (last one shorter), lb = 0, lt ~ N(0, 0.01^2)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    lt = rng.normal(0.0, 0.01, n).astype(np.float32)
    lb = np.zeros(n, dtype=np.float32)
    lens = [traj_len] * (n // traj_len)
    if n % traj_len:
        lens.append(n % traj_len)
    return x, lt, lb, np.asarray(lens, dtype=np.int64)


def constants(d, k):
    import scipy.special
    return float(scipy.special.gamma(d / 2 + 1)), float(math.log(k) - scipy.special.digamma(k))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p["bf16_tflops_sustained"]), "tflops_burst": float(p["bf16_tflops"]),
                "hbm": float(p["hbm_gbs"]), "src": "measured"}
    except Exception:
        return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


# ---------------------------------------------------------------------------------------- CPU arm
def cpu_reference_once(x, lt, lb, lens, k, n_query_rows, G, B):
    """Reference CPU path on a bounded sample: sklearn brute/kd-tree search of `n_query_rows` query
    rows against the full fitted set (cost per query row is independent of how many are asked), and
    the float64 objective + gradient on all N states.  Returns (t_search_sample, t_iw)."""
    from oracle import iw_oracle, knn_oracle
    n = x.shape[0]
    t0 = time.perf_counter()
    dist_s, idx_s = knn_oracle.knn_sklearn(x, k + 1, queries=x[:n_query_rows], n_jobs=os.cpu_count())
    t_search = time.perf_counter() - t0
    # objective on all N states needs full-size neighbour arrays: tile the sampled rows
    reps = (n + n_query_rows - 1) // n_query_rows
    idx = np.tile(idx_s, (reps, 1))[:n]
    dist = np.tile(dist_s, (reps, 1))[:n]
    t0 = time.perf_counter()
    d = x.shape[1]
    iw_oracle.entropy(lt, lb, lens, dist, idx, k, G, B, d, 0.0)
    iw_oracle.kl(lt, lb, lens, idx, k, 0.0)
    iw_oracle.entropy_grad_logp(lt, lb, lens, dist, idx, k, G, d, 0.0)
    t_iw = time.perf_counter() - t0
    return t_search, t_iw


def cpu_sample_rows(x, k, target_s):
    from oracle import knn_oracle
    n = x.shape[0]
    probe = min(n, 512)
    t0 = time.perf_counter()
    knn_oracle.knn_sklearn(x, k + 1, queries=x[:probe], n_jobs=os.cpu_count())
    t = max(time.perf_counter() - t0, 1e-3)
    rows = int(min(n, max(probe, target_s * probe / t)))
    # keep sklearn on the same chunking strategy as a full self-query (parallel_on_X needs
    # n_queries >= 4 * 256 * n_threads, _base.pyx.tp:147)
    rows = max(rows, min(n, 1024 * (os.cpu_count() or 1)))
    return max(256, rows // 256 * 256)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x, lt, lb, lens = synth(args.n, args.d, args.traj_len)
    G, B = constants(args.d, args.k)
    rows = cpu_sample_rows(x, args.k, max(2.0, args.cpu_seconds / 2))
    times = []
    for i in range(args.warmup + args.steps):
        ts, ti = cpu_reference_once(x, lt, lb, lens, args.k, rows, G, B)
        if i >= args.warmup:
            times.append(ts * args.n / rows + ti)
    per_step = float(np.mean(times))
    value = args.n / per_step
    sample = (f"sklearn NearestNeighbors (mepol.py:190-192, algorithm=auto, n_jobs={os.cpu_count()}) on {rows} query "
              f"rows vs all {args.n} fitted rows, scaled x{args.n / rows:.1f}; float64 numpy port of "
              f"compute_entropy+compute_kl+backward on all {args.n} states")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"synthetic sweep N={args.n} d={args.d} k={args.k} L={args.traj_len}"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mepol_b200
    from mepol_b200 import _lib
    from mepol_b200 import entropy as E
    from mepol_b200.dist import ShardedSearch, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise _lib.MepolError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, k = args.n, args.d, args.k
    k1 = k + 1
    if n % world:
        n -= n % world
    algo = {"auto": _lib.KNN_AUTO, "simt": _lib.KNN_SIMT, "tc": _lib.KNN_TC}[args.algo]
    x, lt, lb, lens = synth(n, d, args.traj_len)
    G, B = constants(d, k)
    lo, hi = shard_bounds(n, world, rank)
    n_traj = len(lens)

    # ---- resident inputs (device-timed arm) and pinned host inputs (e2e arm)
    x_pin = torch.from_numpy(x[lo:hi]).pin_memory()
    lt_pin = torch.from_numpy(lt).pin_memory()
    lb_pin = torch.from_numpy(lb).pin_memory()
    shard_dev = x_pin.to(dev)
    lt_dev = lt_pin.to(dev).requires_grad_(True)
    lb_dev = lb_pin.to(dev)
    off_dev, _ = mepol_b200.traj_offsets(lens, dev)
    searcher = ShardedSearch(n, d, k1, dev, algo=algo)
    idx_host = torch.empty((n, k1), dtype=torch.int64).pin_memory() if rank == 0 else None
    dist_host = torch.empty((n, k1), dtype=torch.float64).pin_memory() if rank == 0 else None
    grad_host = torch.empty((n,), dtype=torch.float32).pin_memory() if rank == 0 else None
    scal_host = torch.empty((2,), dtype=torch.float64).pin_memory() if rank == 0 else None

    def objective(idx_full, rk_full, lt_t, lb_t):
        handle = E.SearchHandle(idx_full, rk_full, k1)
        handle.csr()
        H, kl_raw, _ = mepol_b200.iw_entropy(lt_t, lb_t, handle, off_dev, n_traj, k, d, G, B, 0.0)
        lt_t.grad = None
        H.backward()
        return H, kl_raw, lt_t.grad

    def step_device():
        idx_full, rk_full, _ = searcher(shard_dev)
        if rank == 0:
            return objective(idx_full, rk_full, lt_dev, lb_dev)
        return None

    def step_e2e():
        sh = x_pin.to(dev, non_blocking=True)                                   # H2D particles
        idx_full, rk_full, dist_full = searcher(sh, gather_distances=True)
        if rank != 0:
            return None
        idx_host.copy_(idx_full.to(torch.int64), non_blocking=True)             # D2H, sklearn's dtypes
        dist_host.copy_(dist_full, non_blocking=True)
        ltt = lt_pin.to(dev, non_blocking=True).requires_grad_(True)             # H2D log-probs
        lbt = lb_pin.to(dev, non_blocking=True)
        H, kl_raw, g = objective(idx_full, rk_full, ltt, lbt)
        grad_host.copy_(g, non_blocking=True)                                   # D2H gradient
        scal_host.copy_(torch.stack([H.detach(), kl_raw]), non_blocking=True)   # D2H entropy, KL
        torch.cuda.current_stream().synchronize()
        return scal_host

    h2d = (hi - lo) * d * 4 + (2 * n * 4 if rank == 0 else 0)
    d2h = n * k1 * 16 + n * 4 + 16

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.profile_enable(True)
    ms_total = timed(step_device, args.steps)
    prune_ms = _lib.profile_read(64)
    _lib.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = n / (ms_per_step * 1e-3)

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps) / args.steps
        e2e = {"value": n / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e}

    # ---- component timings (outside the timed region; explanatory only)
    comp = {}
    if rank == 0:
        def t_of(fn, reps=3):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        idx_full, rk_full, _ = searcher(shard_dev) if world == 1 else (searcher.idx_full, searcher.rk_full, None)
        handle = E.SearchHandle(idx_full, rk_full, k1)
        comp["rev_csr_ms"] = t_of(lambda: E.SearchHandle(idx_full, rk_full, k1).csr())
        handle.csr()

        def fb():
            H, _, _ = mepol_b200.iw_entropy(lt_dev, lb_dev, handle, off_dev, n_traj, k, d, G, B, 0.0)
            lt_dev.grad = None
            H.backward()
        comp["iw_fwd_bwd_ms"] = t_of(fb, 10)
        # the kernels alone, straight through the C ABI with pre-allocated buffers (no autograd / allocator)
        lib = _lib.load()
        c_b = torch.empty(n, dtype=torch.float32, device=dev)
        gw_b = torch.empty(n, dtype=torch.float32, device=dev)
        gr_b = torch.empty(n, dtype=torch.float32, device=dev)
        sc_b = torch.empty(8, dtype=torch.float64, device=dev)
        one_b = torch.ones(1, dtype=torch.float64, device=dev)
        ws_b = torch.empty(_lib.iw_workspace_bytes(n, n_traj), dtype=torch.uint8, device=dev)
        rowptr, col = handle.csr()
        rk32 = handle.rk(torch.float32)
        sp = torch.cuda.current_stream().cuda_stream
        ltd = lt_dev.detach()

        def k_fwd():
            _lib.check(lib.mepol_iw_entropy_fwd_f32(ltd.data_ptr(), lb_dev.data_ptr(), off_dev.data_ptr(), n_traj, n,
                                                    handle.idx.data_ptr(), k1, rk32.data_ptr(), d, math.log(G), B, 0.0,
                                                    c_b.data_ptr(), gw_b.data_ptr(), sc_b.data_ptr(), ws_b.data_ptr(),
                                                    ws_b.numel(), sp), "fwd")

        def k_bwd():
            _lib.check(lib.mepol_iw_entropy_bwd_f32(c_b.data_ptr(), gw_b.data_ptr(), sc_b.data_ptr(), off_dev.data_ptr(),
                                                    n_traj, n, rowptr.data_ptr(), col.data_ptr(), one_b.data_ptr(),
                                                    gr_b.data_ptr(), sp), "bwd")
        k_fwd()
        comp["iw_kernels_fwd_ms"] = t_of(k_fwd, 20)
        comp["iw_kernels_bwd_ms"] = t_of(k_bwd, 20)
        if world == 1:
            comp["search_ms"] = t_of(lambda: searcher(shard_dev), 2)
            comp["search_stats"] = {k_: int(v) for k_, v in zip(("candidates", "fallback_rows", "compactions"),
                                                               searcher._out[3].cpu().tolist()[:3])}
            # the epoch-realistic mix of the reference's off-policy loop -- one search per epoch and
            # max_off_iters = 30 objective evaluations (forward + backward) on the same particles
            comp["epoch_mix_states_per_s"] = n / ((comp["search_ms"] + 30.0 * comp["iw_fwd_bwd_ms"]) * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    prune_avg = float(np.mean(prune_ms)) if prune_ms else float("nan")
    flops = 2.0 * (hi - lo) * n * d
    achieved = flops / (prune_avg * 1e-3) / 1e12 if prune_ms else None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["tflops"]) if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
                # (profiles/r01_ncu_tc_v10_raw.txt); only valid for the configuration that capture was taken on
                "traffic": 1.610e9 if (n == 1048576 and world == 1 and d == 29 and algo != _lib.KNN_SIMT) else None,
                "kernel": "knn_prune_tc" if (algo != _lib.KNN_SIMT and prune_avg < 400) else "knn_prune",
                "kernel_ms": prune_avg, "launches_timed": len(prune_ms),
                "algorithmic_flops_per_launch": flops,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})"}
    comp["iw_roofline"] = None
    if "iw_kernels_fwd_ms" in comp:
        bytes_iw = (36 + 20 * k) * n
        t_k = comp["iw_kernels_fwd_ms"] + comp["iw_kernels_bwd_ms"]
        gbs = bytes_iw / (t_k * 1e-3) / 1e9
        comp["iw_roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
                               "frac": gbs / peaks["hbm"], "algorithmic_bytes": bytes_iw, "kernel_ms": t_k,
                               "note": "IW fwd (scan+gather) + bwd kernels through the C ABI; (36+20k) B/state"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rows = cpu_sample_rows(x, k, args.cpu_seconds)
        ts, ti = cpu_reference_once(x, lt.astype(np.float64), lb.astype(np.float64), lens, k, rows, G, B)
        v = n / (ts * n / rows + ti)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": (f"sklearn NearestNeighbors(algorithm=auto, n_jobs={os.cpu_count()}) exactly as mepol.py:190-192 on "
                          f"{rows} query rows vs all {n} fitted rows ({ts:.1f} s, scaled x{n / rows:.1f}) + float64 numpy "
                          f"port of compute_entropy/compute_kl/backward on all {n} states ({ti:.2f} s)")}

    # counted in profiles/r01_ncu_launches_bench_v10.csv -- search: column sums + final, pack x2, prune, rescue,
    # re-rank, fallback list/partial/merge/serial = 11; reverse CSR 6; IW scan, gather, backward elem + reverse scan = 4
    launches_per_step = 11 + 6 + 4
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if algo == _lib.KNN_SIMT else "f16",
            "data": "synthetic",
            "config": {"workload": f"synthetic sweep N={n} d={d} k={k} L={args.traj_len} (BASELINE.json configs[4])",
                       "algo": args.algo, "parallelism": f"query rows sharded x{world}, refs all-gathered (NCCL)",
                       "l2": "inputs larger than L2: packed refs + candidates > 126 MB per step, no flush needed"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "components": comp}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
