"""Benchmark of the hot path: exact k-NN search + IW entropy forward/backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): k-NN entropy states/sec (search + IW fwd/bwd) at N=1M, d=29, k=4.
One "step" = one search of all N synthetic particles (k+1=5, self included), the reverse-CSR build
for that search, and one IW-entropy forward (H + KL) + backward on the same N states.
Prints ONE JSON line on rank 0.
"""
import os
import sys

# torchrun exports OMP_NUM_THREADS=1 when it is unset; the reference arm is a CPU measurement "with all
# the host threads it can use", so undo that BEFORE numpy / sklearn / torch load their OpenMP runtimes.
if "reference" in sys.argv and os.environ.get("OMP_NUM_THREADS") == "1" and (
        "TORCHELASTIC_RUN_ID" in os.environ or "LOCAL_RANK" in os.environ):
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import argparse
import json
import math
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "k-NN entropy states/sec (search + IW fwd/bwd)"
UNIT = "states/s"
# kernels of libmepol_b200.so by name prefix (for the live launch count)
OUR_KERNEL_PREFIXES = ("knn_", "iw_", "csr_", "pack_rows", "col_sum", "col_mean", "grid_", "rerank", "fallback",
                       "widen_idx", "merge_", "rescue")


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
    ap.add_argument("--workload", default="sweep", choices=["sweep"] + sorted(CONFIG_WORKLOADS),
                    help="sweep = BASELINE.json configs[4] (the metric's configuration, --n/--d/--k apply); the others are "
                         "the reference's shipped configurations (configs[0..3]) as synthetic batches of their shapes")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU-baseline sample time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the result checks (parity key)")
    return ap.parse_args()


# BASELINE.json configs[0..3] as synthetic batches (rollouts need gym / MuJoCo):
#   name -> (num_traj T, traj_len L, ns = particle dimension, k, particle generator)
CONFIG_WORKLOADS = {
    "grid_world": (20, 1200, 2, 50, "uniform6"),        # scripts/tae/grid_world.sh (k = 50 per the script)
    "grid_world_k4": (20, 1200, 2, 4, "uniform6"),      # BASELINE.json's k = 4 reading of the same config
    "mountain_car": (20, 400, 2, 4, "mountaincar"),     # scripts/tae/mountain_car.sh
    "ant": (20, 500, 7, 4, "normal"),                   # scripts/tae/ant.sh, state_filter = range(7)
    "ant_29d": (20, 500, 29, 4, "normal"),              # BASELINE.json's "29-dim" reading
    "humanoid": (20, 500, 24, 4, "normal"),             # scripts/tae/humanoid.sh, ns = 24
}


def apply_workload(args):
    if args.workload != "sweep":
        T, L, ns, k, _ = CONFIG_WORKLOADS[args.workload]
        args.n, args.d, args.k, args.traj_len = T * L, ns, k, L


def workload_name(n, d, k, traj_len, workload="sweep"):
    """ONE string for both arms (the driver compares the two lines' config)."""
    if workload != "sweep":
        return f"{workload}: synthetic batch of the shipped configuration's shape N={n} ns={d} k={k} L={traj_len}"
    return f"synthetic sweep N={n} d={d} k={k} L={traj_len} (BASELINE.json configs[4])"


def synth(n, d, traj_len, seed=0, kind="normal"):
    """Synthetic batch of BASELINE.md section 4: N(0,1) float32 particles (tie-free), trajectories of
    length traj_len (last one shorter), lb = 0, lt ~ N(0, 0.01^2).  kind: the particle generator of the
    shipped configurations (GridWorld uniform in [-6,6]^2, MountainCar x in [-1.2,0.6], v in [-0.07,0.07])."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    if kind == "uniform6":
        x = rng.uniform(-6.0, 6.0, (n, d)).astype(np.float32)
    elif kind == "mountaincar":
        x = np.stack([rng.uniform(-1.2, 0.6, n), rng.uniform(-0.07, 0.07, n)], 1).astype(np.float32)
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


def traffic_from_capture(kernel, n, d, world):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture, but only
    when that capture was taken on this kernel and this configuration (profiles/*_traffic.json); else null."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                t = json.load(f)
            if t.get("kernel") == kernel and t.get("n") == n and t.get("d") == d and t.get("n_gpus", 1) == world:
                return float(t["dram_bytes_per_launch"]), os.path.basename(path)
        except Exception:
            continue
    return None, None


# ---------------------------------------------------------------------------------------- CPU arm
def host_threads():
    """Use every host core (and say how many the thread pools really have)."""
    n = os.cpu_count() or 1
    info = []
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(n)
        info = [{k: v for k, v in lib.items() if k in ("user_api", "internal_api", "num_threads", "version")}
                for lib in threadpoolctl.threadpool_info()]
    except Exception:
        pass
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    return n, info


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


class CpuReference:
    """The reference's CPU implementation of the path on a bounded sample of the workload.

    search   : sklearn's NearestNeighbors exactly as mepol.py:190-192 (fit on all N rows, kneighbors on
               `rows` of them; the cost per query row of a brute search does not depend on how many rows
               are asked, so the time is scaled by N / rows and labelled as such);
    objective: when oracle/_ref is staged (oracle/make_ref.py) the UNMODIFIED reference functions
               compute_entropy + loss.backward() + compute_kl (mepol.py:142-174, :278) with a float64
               GaussianPolicy([400, 300], 29, 8) (policy.py:11-61) on `n_obj` states (whole trajectories),
               scaled by N / n_obj -> kind "reference"; otherwise the numpy port (oracle/iw_oracle.py)
               on all N states -> kind "port"."""

    def __init__(self, x, lt, lb, lens, k, traj_len, obj_states=131072):
        self.x, self.lt, self.lb, self.lens, self.k = x, lt, lb, lens, k
        self.n, self.d = x.shape
        self.G, self.B = constants(self.d, k)
        self.kind = "port"
        self.ref = None
        try:
            from oracle import make_ref
            if make_ref.available():
                import torch
                prev = torch.get_default_dtype()
                self.ref, ref_policy = make_ref.import_reference()     # flips the default dtype (dtypes.py:3)
                self.torch = torch
                T = max(1, min(len(lens), obj_states // traj_len))
                while T > 1 and lens[T - 1] != traj_len:
                    T -= 1
                self.T, self.L = T, int(lens[0]) if lens[0] == traj_len else int(lens[0])
                self.n_obj = int(np.sum(lens[:T]))
                rng = np.random.default_rng(1)
                A = 8
                torch.manual_seed(0)
                self.behav = ref_policy.GaussianPolicy(num_features=self.d, hidden_sizes=[400, 300], action_dim=A)
                self.target = ref_policy.GaussianPolicy(num_features=self.d, hidden_sizes=[400, 300], action_dim=A)
                self.target.load_state_dict(self.behav.state_dict())
                with torch.no_grad():
                    for p in self.target.parameters():
                        p.add_(1e-3 * torch.randn_like(p))
                self.states = torch.tensor(rng.standard_normal((T, self.L + 1, self.d)), dtype=torch.float64)
                self.actions = torch.tensor(0.6 * rng.standard_normal((T, self.L, A)), dtype=torch.float64)
                self.rtl = torch.full((T, 1), self.L, dtype=torch.int64)
                self.kind = "reference"
                self._restore = prev
        except Exception as e:        # the port keeps the arm alive; the line says which one ran
            self.ref = None
            self.kind = "port"
            self.why_port = repr(e)

    def search(self, rows):
        from oracle import knn_oracle
        t0 = time.perf_counter()
        dist_s, idx_s = knn_oracle.knn_sklearn(self.x, self.k + 1, queries=self.x[:rows], n_jobs=os.cpu_count())
        return time.perf_counter() - t0, dist_s, idx_s

    def objective(self, dist_s, idx_s):
        """Seconds for one objective evaluation (entropy + backward + KL), scaled to all N states."""
        rows = idx_s.shape[0]
        if self.ref is not None:
            torch = self.torch
            n_obj = self.n_obj
            reps = (n_obj + rows - 1) // rows
            # neighbour lists of the sampled rows tiled over the objective sample (valid indices; the
            # reference's arithmetic does not depend on which neighbours they are)
            idx = torch.tensor(np.tile(idx_s % n_obj, (reps, 1))[:n_obj], dtype=torch.int64)
            dist = torch.tensor(np.tile(dist_s, (reps, 1))[:n_obj], dtype=torch.float64)
            t0 = time.perf_counter()
            for p in self.target.parameters():
                p.grad = None
            ent = self.ref.compute_entropy(self.behav, self.target, self.states, self.actions, self.T, self.rtl,
                                           dist, idx, self.k, self.G, self.B, self.d, 0.0)
            (-ent).backward()
            with torch.no_grad():
                self.ref.compute_kl(self.behav, self.target, self.states, self.actions, self.T, self.rtl,
                                    dist, idx, self.k, 0.0)
            return (time.perf_counter() - t0) * self.n / n_obj
        from oracle import iw_oracle
        reps = (self.n + rows - 1) // rows
        idx = np.tile(idx_s, (reps, 1))[:self.n]
        dist = np.tile(dist_s, (reps, 1))[:self.n]
        lt64, lb64 = self.lt.astype(np.float64), self.lb.astype(np.float64)
        t0 = time.perf_counter()
        iw_oracle.entropy(lt64, lb64, self.lens, dist, idx, self.k, self.G, self.B, self.d, 0.0)
        iw_oracle.kl(lt64, lb64, self.lens, idx, self.k, 0.0)
        iw_oracle.entropy_grad_logp(lt64, lb64, self.lens, dist, idx, self.k, self.G, self.d, 0.0)
        return time.perf_counter() - t0

    def describe(self, rows, ts, ti, cores):
        s = (f"sklearn NearestNeighbors(algorithm=auto, n_jobs={cores}) exactly as mepol.py:190-192: fit on all {self.n} "
             f"rows, kneighbors on {rows} of them ({ts:.2f} s, scaled x{self.n / rows:.1f}); ")
        if self.ref is not None:
            s += (f"unmodified reference compute_entropy + backward + compute_kl (mepol.py:142-174,:278, float64, "
                  f"GaussianPolicy([400,300],{self.d},8)) from oracle/_ref on {self.n_obj} states "
                  f"({ti * self.n_obj / self.n:.2f} s, scaled x{self.n / self.n_obj:.1f})")
        else:
            s += f"float64 numpy port of compute_entropy/compute_kl/backward on all {self.n} states ({ti:.2f} s)"
        return s

    def versions(self):
        import sklearn
        import scipy
        import torch
        return {"sklearn": sklearn.__version__, "torch": torch.__version__, "numpy": np.__version__,
                "scipy": scipy.__version__}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores, pools = host_threads()
    kind = CONFIG_WORKLOADS[args.workload][4] if args.workload != "sweep" else "normal"
    x, lt, lb, lens = synth(args.n, args.d, args.traj_len, args.seed, kind)
    ref = CpuReference(x, lt, lb, lens, args.k, args.traj_len)
    rows = cpu_sample_rows(x, args.k, max(2.0, args.cpu_seconds / 2))
    warm = min(args.warmup, 1)
    steps = max(1, min(args.steps, 3))              # each step is seconds of CPU work; more repeats buy nothing
    times, ts, ti = [], 0.0, 0.0
    for i in range(warm + steps):
        ts, dist_s, idx_s = ref.search(rows)
        ti = ref.objective(dist_s, idx_s)
        if i >= warm:
            times.append(ts * args.n / rows + ti)
    per_step = float(np.mean(times))
    value = args.n / per_step
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.n, args.d, args.k, args.traj_len, args.workload), "seed": args.seed},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": ref.kind,
                             "sample": ref.describe(rows, ts, ti, cores), "threadpools": pools,
                             "versions": ref.versions(), "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "steps_requested": args.steps, "warmup_requested": args.warmup}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------- GPU arm
def count_our_launches(fn):
    """Launches of this repo's kernels in ONE step, counted live with torch.profiler (CUPTI).  Returns
    (count, {kernel name: launches}) or (None, reason)."""
    try:
        import torch
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        names = {}
        for ev in prof.events():
            nm = ev.name
            if any(p in nm for p in OUR_KERNEL_PREFIXES) and "Memcpy" not in nm and "Memset" not in nm:
                short = nm.split("<")[0].split("(")[0].replace("void ", "").replace("mepol::", "")
                names[short] = names.get(short, 0) + 1
        total = sum(names.values())
        if total == 0:
            return None, "profiler saw none of this library's kernels"
        return total, names
    except Exception as e:
        return None, repr(e)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import mepol_b200
    from mepol_b200 import _lib
    from mepol_b200 import entropy as E
    from mepol_b200.dist import ShardedSearch, SharedHostResults, shard_bounds

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
    kind = CONFIG_WORKLOADS[args.workload][4] if args.workload != "sweep" else "normal"
    x, lt, lb, lens = synth(n, d, args.traj_len, args.seed, kind)
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
    # sklearn's output arrays on the host (indices int64, distances float64): one page-locked shared-memory
    # block that every rank's GPU writes its own rows into (G concurrent PCIe transfers)
    host_results, host_results_note = None, None
    if not args.no_e2e:
        try:
            host_results = SharedHostResults(n, k1, device=dev)
        except Exception as e:          # e.g. a tiny /dev/shm: fall back to gathering on rank 0's GPU (the slower path)
            host_results_note = f"shared host block unavailable ({e!r}); results gathered on rank 0's GPU instead"
            fb_idx = torch.empty((n, k1), dtype=torch.int64).pin_memory() if rank == 0 else None
            fb_dist = torch.empty((n, k1), dtype=torch.float64).pin_memory() if rank == 0 else None
    grad_host = torch.empty((n,), dtype=torch.float32).pin_memory() if rank == 0 else None
    scal_host = torch.empty((2,), dtype=torch.float64).pin_memory() if rank == 0 else None
    copy_stream = torch.cuda.Stream(device=dev)

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
        sh = x_pin.to(dev, non_blocking=True)                                   # H2D particles (own shard)
        if rank == 0:
            # H2D of the log-probs on the copy stream: it overlaps the search instead of sitting between the search
            # and the objective (the tensors stay referenced until the step's final synchronize)
            with torch.cuda.stream(copy_stream):
                lt_in = lt_pin.to(dev, non_blocking=True)
                lbt = lb_pin.to(dev, non_blocking=True)
                logp_in = torch.cuda.Event()
                logp_in.record(copy_stream)
        idx_full, rk_full, dist_full = searcher(sh, gather_distances=host_results is None)
        # every rank sends the rows it searched to the host arrays on a copy stream (D2H indices as int64 +
        # distances float64: 16 * k1 bytes per row), while rank 0 goes on with the objective
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            if host_results is not None:
                host_results.push(searcher.own_idx, searcher.own_dist)
            elif rank == 0:
                fb_idx.copy_(idx_full, non_blocking=True)
                fb_dist.copy_(dist_full, non_blocking=True)
        if rank == 0:
            torch.cuda.current_stream().wait_event(logp_in)
            ltt = lt_in.requires_grad_(True)
            H, kl_raw, g = objective(idx_full, rk_full, ltt, lbt)
            grad_host.copy_(g, non_blocking=True)                               # D2H gradient
            scal_host.copy_(torch.stack([H.detach(), kl_raw]), non_blocking=True)   # D2H entropy, KL
        copy_stream.synchronize()
        torch.cuda.current_stream().synchronize()
        if world > 1:
            dist.barrier()                                                      # all rows have landed on the host
        return scal_host

    h2d = n * d * 4 + 2 * n * 4                     # all ranks' shards + the log-probs on rank 0
    d2h = n * k1 * 16 + n * 4 + 16                  # all ranks' result rows + gradient and scalars on rank 0

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

    warmup_done = max(args.warmup, 3)
    for _ in range(warmup_done):
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

    # ---- per-phase breakdown of one step on this rank, max over ranks (outside the timed region)
    phases = searcher.timed_phases(shard_dev)
    if world > 1:
        keys = sorted(phases)
        t = torch.tensor([phases[k_] for k_ in keys], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = dict(zip(keys, [float(v) for v in t.tolist()]))

    # ---- result checks (parity key)
    parity = None
    if not args.no_verify:
        idx_full, rk_full, dist_full = searcher(shard_dev, gather_distances=True)
        if host_results is not None and rank == 0:
            # the e2e arm's host arrays (written by all ranks in the last e2e step) against the device result
            host_ok = bool((host_results.indices == idx_full.cpu().to(torch.int64)).all()) and \
                bool((host_results.distances == dist_full.cpu()).all())
        else:
            host_ok = None
        if world > 1 and rank == 0:
            # the same array searched by ONE GPU (rank 0, all rows): the sharded result must be bit-identical
            full_dev = searcher.full
            single = mepol_b200.knn_device(full_dev, k1, algo=algo)
            mism = int((single.indices != idx_full).any(dim=1).sum().item())
            mism_rk = int((single.rk != rk_full).sum().item())
            parity = {"vs": "single-GPU search of the same array on rank 0 (bitwise)", "rows_checked": n,
                      "mismatch": mism, "rk_mismatch": mism_rk, "host_arrays_equal_device_result": host_ok}
            del single
        elif world == 1:
            parity = {"vs": None, "rows_checked": 0, "mismatch": None, "host_arrays_equal_device_result": host_ok}
        if world > 1:
            dist.barrier()

    # ---- component timings (outside the timed region; explanatory only)
    comp = {"phases_ms": phases}
    launches, launch_names = None, None
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
        idx_full, rk_full = searcher.last_idx, searcher.last_rk
        handle = E.SearchHandle(idx_full, rk_full, k1)
        comp["rev_csr_ms"] = t_of(lambda: E.SearchHandle(idx_full, rk_full, k1).csr())
        handle.csr()

        def fb():
            H, _, _ = mepol_b200.iw_entropy(lt_dev, lb_dev, handle, off_dev, n_traj, k, d, G, B, 0.0)
            lt_dev.grad = None
            H.backward()
        comp["iw_fwd_bwd_ms"] = t_of(fb, 10)
        # the same forward + backward replayed as ONE CUDA graph -- what entropy.policy_update does in production (the
        # eager autograd call is host-bound: ~0.2-0.35 ms of Python per evaluation around 0.06 ms of kernels)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fb()
                fb()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fb()
            graph.replay()
            comp["iw_fwd_bwd_graph_ms"] = t_of(graph.replay, 20)
        except Exception as e:      # explanatory number only
            comp["iw_fwd_bwd_graph_ms"] = None
            comp["iw_fwd_bwd_graph_error"] = repr(e)[:200]
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
        def k_fused():
            _lib.check(lib.mepol_iw_entropy_fwd_bwd_f32(ltd.data_ptr(), lb_dev.data_ptr(), off_dev.data_ptr(), n_traj, n,
                                                        handle.idx.data_ptr(), k1, rk32.data_ptr(), d, math.log(G), B, 0.0,
                                                        rowptr.data_ptr(), col.data_ptr(), c_b.data_ptr(), gw_b.data_ptr(),
                                                        sc_b.data_ptr(), gr_b.data_ptr(), ws_b.data_ptr(), ws_b.numel(), sp),
                       "fwd_bwd")
        k_fwd()
        comp["iw_kernels_fwd_ms"] = t_of(k_fwd, 20)
        comp["iw_kernels_bwd_ms"] = t_of(k_bwd, 20)
        # one C-ABI call for forward + backward (what the autograd bridge issues), buffers pre-allocated
        comp["iw_fused_call_ms"] = t_of(k_fused, 20)
        if world == 1:
            comp["search_ms"] = t_of(lambda: searcher(shard_dev), 2)
            comp["search_stats"] = {k_: int(v) for k_, v in zip(("candidates", "fallback_rows", "compactions"),
                                                               searcher._out[3].cpu().tolist()[:3])}
            # the epoch-realistic mix of the reference's off-policy loop -- one search per epoch and
            # max_off_iters = 30 objective evaluations (forward + backward) on the same particles
            comp["epoch_mix_states_per_s"] = n / ((comp["search_ms"] + 30.0 * comp["iw_fwd_bwd_ms"]) * 1e-3)
            comp["epoch_mix_states_per_s_c_abi"] = n / ((comp["search_ms"] + 30.0 * comp["iw_fused_call_ms"]) * 1e-3)
            if comp.get("iw_fwd_bwd_graph_ms"):
                comp["epoch_mix_states_per_s_graph"] = n / ((comp["search_ms"] + 30.0 * comp["iw_fwd_bwd_graph_ms"]) * 1e-3)
    # live launch count of one step (rank 0 launches the objective as well as its share of the search)
    if world > 1:
        dist.barrier()
    if rank == 0 or world > 1:
        lc, ln = count_our_launches(step_device)
        if rank == 0:
            launches, launch_names = lc, ln
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    prune_avg = float(np.mean(prune_ms)) if prune_ms else float("nan")
    flops = 2.0 * (hi - lo) * n * d
    achieved = flops / (prune_avg * 1e-3) / 1e12 if prune_ms else None
    kernel = _lib.prune_kernel_name(algo, d, k1, n)
    traffic, traffic_src = traffic_from_capture(kernel, n, d, world)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": (achieved / peaks["tflops"]) if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture
                # of THIS kernel on THIS configuration (null when there is none)
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": kernel, "kernel_ms": prune_avg, "launches_timed": len(prune_ms),
                "algorithmic_flops_per_launch": flops,
                "peak_burst": peaks["tflops_burst"],
                "frac_of_burst": (achieved / peaks["tflops_burst"]) if achieved else None,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']}): the kernel is timed inside a long step"}
    if kernel == "knn_grid_search_kernel":
        # d <= 3 (the reference's kd-tree regime): O(N k) uniform-grid search, bounded by memory not by the tensor pipe
        gbytes = float((hi - lo) * (4 * d + 16 + k1 * 12 + 8))
        gbs = gbytes / (prune_avg * 1e-3) / 1e9 if prune_ms else None
        roofline = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": (gbs / peaks["hbm"]) if gbs else None, "traffic": None, "kernel": kernel, "kernel_ms": prune_avg,
                    "launches_timed": len(prune_ms), "algorithmic_bytes_per_launch": gbytes,
                    "note": "particles in + 16 B grid point + (k+1) x (int32 index + float64 distance) + r_k out per query row; "
                            "latency-bound ring walk, the figure is a lower bound of the traffic",
                    "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peaks['src']})"}
    comp["iw_roofline"] = None
    if "iw_kernels_fwd_ms" in comp:
        bytes_iw = (36 + 20 * k) * n
        t_k = comp["iw_kernels_fwd_ms"] + comp["iw_kernels_bwd_ms"]
        gbs = bytes_iw / (t_k * 1e-3) / 1e9
        # What actually paces the two gather kernels: every neighbour weight / reverse-neighbour gradient is a random
        # 4-byte read of an L2-resident array (4 MB at N = 1M) and moves a whole 32-byte sector to the SM: 2 k N sectors
        # per forward + backward on top of the streamed bytes (ncu: DRAM reads below the algorithmic bytes, 0 B written,
        # profiles/r02_ncu_iw_kernels.txt).  Stated against the L2 throughput cap of the profiling guide
        # (~6300 B/clk full chip) at the SM clock sampled during the timed region.
        sector_bytes = 2 * k * n * 32
        l2_cap = 6300.0 * (clocks["sm_mhz"] if clocks and clocks.get("sm_mhz") else 1900.0) * 1e6 / 1e9
        comp["iw_roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s",
                               "frac": gbs / peaks["hbm"], "algorithmic_bytes": bytes_iw, "kernel_ms": t_k,
                               "note": "IW fwd (scan+gather) + bwd kernels through the C ABI; (36+20k) B/state",
                               "l2": {"sector_bytes_random_gathers": sector_bytes,
                                      "achieved_gbs": (bytes_iw + sector_bytes) / (t_k * 1e-3) / 1e9,
                                      "cap_gbs": l2_cap, "frac": (bytes_iw + sector_bytes) / (t_k * 1e-3) / 1e9 / l2_cap,
                                      "note": "streamed bytes + 32-byte sectors of the 2kN random 4-byte gathers; the working "
                                              "set is L2-resident, so this (not HBM) is the bound of the gather kernels"}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores, pools = host_threads()
        ref = CpuReference(x, lt, lb, lens, k, args.traj_len)
        rows = cpu_sample_rows(x, k, args.cpu_seconds)
        ts, dist_s, idx_s = ref.search(rows)
        ti = ref.objective(dist_s, idx_s)
        if getattr(ref, "_restore", None) is not None:
            torch.set_default_dtype(ref._restore)
        v = n / (ts * n / rows + ti)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": ref.describe(rows, ts, ti, cores),
               "threadpools": pools, "versions": ref.versions()}
        if not args.no_verify:
            # the benchmarked array itself: the GPU's rows against sklearn's answer for the sampled rows
            got_i = searcher.last_idx[:rows].cpu().numpy().astype(np.int64)
            got_d = searcher.last_dist[:rows].cpu().numpy()
            bad_rows = int((got_i != idx_s).any(axis=1).sum())
            far = dist_s > 1e-3                                         # col 0 (self) is ~1e-7 in sklearn's brute form
            rel = float(np.max(np.abs(got_d[far] - dist_s[far]) / dist_s[far]))
            parity = {"vs": "sklearn NearestNeighbors on the benchmarked array (the CPU leg's sampled rows)",
                      "rows_checked": int(rows), "mismatch": bad_rows, "max_rel_dist_err": rel, "dist_rtol": 1e-5,
                      "host_arrays_equal_device_result": host_ok}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup_done, "warmup_requested": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if algo == _lib.KNN_SIMT else "f16",
            "data": "synthetic",
            "config": {"workload": workload_name(n, d, k, args.traj_len, args.workload), "seed": args.seed,
                       "algo": args.algo, "parallelism": f"query rows sharded x{world}, refs all-gathered (NCCL)",
                       "l2": "inputs larger than L2: packed refs + candidates > 126 MB per step, no flush needed"},
            "clocks": clocks, "e2e": (dict(e2e, note=host_results_note) if (e2e and host_results_note) else e2e),
            "gpu_launches": (launches * args.steps) if launches else None,
            "gpu_launches_per_step": launch_names if launches else {"uncounted": launch_names},
            "parity": parity, "roofline": roofline, "cpu_baseline": cpu, "components": comp}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    apply_workload(a)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
