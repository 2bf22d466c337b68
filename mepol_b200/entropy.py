"""Host side of the importance-weighted k-NN entropy / KL objective.

Mirrors, with the same names, argument order and return conventions, the module-level functions
of /root/reference/src/algorithms/mepol.py that `mepol()` calls:

    compute_importance_weights  :114-139     compute_entropy :142-154
    compute_kl                  :157-174     policy_update   :268-281

Policies stay the reference's CPU float64 `GaussianPolicy` modules (src/policy.py:11-61) and the
optimiser stays a CPU torch optimiser; results come back as CPU 0-dim float64 tensors that support
`.backward()`, `.detach()`, `.numpy()`, `torch.isnan`.
Underneath: parameters are mirrored to the GPU, the MLP log-probs run in torch on the device
(cuBLAS, float64 by default), and everything from the log-probs to the entropy, the KL and the
gradient w.r.t. the log-probs runs in the hand-written kernels of libmepol_b200.so through
`IWEntropyFunction` (a torch.autograd.Function).  No CPU fallback.
"""
import math
import weakref

import numpy as np
import torch

from . import _lib
from .knn import _stream_ptr

POLICY_EPS = 1e-7          # src/utils/dtypes.py:7, used inside get_log_p's variance term only
_LOG_2PI = math.log(2.0 * math.pi)

_config = {"compute_dtype": torch.float64, "burn_rng": True, "device": None, "cuda_graphs": True}


def configure(compute_dtype=None, burn_rng=None, device=None, cuda_graphs=None):
    """compute_dtype: torch.float64 (default, reproduces the float64 reference to ~1e-12) or
    torch.float32 (the bandwidth-optimal kernels; 1e-5 relative).  burn_rng: consume the same
    torch.randn draws the reference's get_log_p burns so seeded runs stay aligned.
    cuda_graphs: replay the device side of an objective step as one captured CUDA graph per
    (batch, search result) instead of ~60 separate launches (the step is launch-bound at the
    reference's batch sizes); same kernels, same results."""
    if compute_dtype is not None:
        if compute_dtype not in (torch.float32, torch.float64):
            raise ValueError("compute_dtype must be torch.float32 or torch.float64")
        _config["compute_dtype"] = compute_dtype
    if burn_rng is not None:
        _config["burn_rng"] = bool(burn_rng)
    if device is not None:
        _config["device"] = torch.device(device)
    if cuda_graphs is not None:
        _config["cuda_graphs"] = bool(cuda_graphs)


def _device():
    if not torch.cuda.is_available():
        raise _lib.MepolError("mepol_b200 needs a CUDA device (there is no CPU fallback)")
    return _config["device"] or torch.device("cuda", torch.cuda.current_device())


import contextlib

_NULL_CTX = contextlib.nullcontext()


def _on(dev):
    """Context that makes `dev` current -- free when it already is (the guard costs ~10 us per call, a
    tenth of the whole objective at small batch sizes)."""
    return _NULL_CTX if dev.index == torch.cuda.current_device() else torch.cuda.device(dev)


# ------------------------------------------------------------------------------- device handles
class SearchHandle:
    """Device-resident view of one k-NN search result, shared by every objective evaluation of an
    epoch: int32 indices, the k-th-neighbour radius column, trajectory offsets and (lazily) the
    reverse-neighbour CSR used by the backward."""

    def __init__(self, idx_i32, rk_f64, k1):
        self.idx = idx_i32.contiguous()
        self.k1 = int(k1)
        self.n = int(idx_i32.shape[0])
        self._rk = {torch.float64: rk_f64.contiguous()}
        self._csr = None
        self._ws = None
        self._scratch = {}

    def rk(self, dtype):
        if dtype not in self._rk:
            self._rk[dtype] = self._rk[torch.float64].to(dtype)
        return self._rk[dtype]

    def scratch(self, dtype, n_traj):
        """(c, gW, workspace) of the objective kernels for this search result: intermediates that never
        leave a forward/backward call, allocated once per (dtype, trajectory count) and reused."""
        key = (dtype, int(n_traj))
        hit = self._scratch.get(key)
        if hit is None:
            dev = self.idx.device
            with torch.cuda.device(dev):
                hit = (torch.empty(self.n, dtype=dtype, device=dev), torch.empty(self.n, dtype=dtype, device=dev),
                       torch.empty(_lib.iw_workspace_bytes(self.n, n_traj), dtype=torch.uint8, device=dev))
            self._scratch = {key: hit}
        return hit

    def csr(self):
        if self._csr is None:
            dev = self.idx.device
            k = self.k1 - 1
            with torch.cuda.device(dev):
                rowptr = torch.empty(self.n + 1, dtype=torch.int32, device=dev)
                col = torch.empty(self.n * k, dtype=torch.int32, device=dev)
                ws = torch.empty(_lib.rev_csr_workspace_bytes(self.n, k), dtype=torch.uint8, device=dev)
                rc = _lib.load().mepol_rev_csr(self.idx.data_ptr(), self.n, self.k1, rowptr.data_ptr(),
                                               col.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
                _lib.check(rc, "mepol_rev_csr")
            self._csr = (rowptr, col)
            self._ws = ws
        return self._csr


def traj_offsets(real_traj_lengths, device):
    lens = torch.as_tensor(real_traj_lengths).reshape(-1).to(torch.int64).cpu()
    off = torch.zeros(lens.numel() + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(lens, 0)
    return off.to(torch.int32).to(device), int(off[-1])


# ------------------------------------------------------------------------------- autograd bridge
class IWEntropyFunction(torch.autograd.Function):
    """(H, KL_raw, n_bad) = f(logp_target, logp_behav) through the fused CUDA kernels.

    forward  -> mepol_iw_entropy_fwd_{f32,f64}       (replaces mepol.py:121-152 and :159-165), or, when a
                gradient will be asked for, mepol_iw_entropy_fwd_bwd_{f32,f64}: forward AND backward for
                dH = 1 in ONE library call (the objective is launch/host bound at every batch size)
    backward -> one multiply by the incoming dH (replaces autograd of the same lines, mepol.py:278)

    Scratch that never leaves the call (c, gW, the reduction workspace) lives on the SearchHandle and is
    reused by every evaluation of the epoch; only what the caller can keep (the scalars, the gradient)
    is allocated per call."""

    @staticmethod
    def forward(ctx, logp_t, logp_b, handle, traj_off, n_traj, k, ns, log_G, B, eps):
        if not logp_t.is_cuda:
            raise _lib.MepolError("IWEntropyFunction needs CUDA tensors (there is no CPU fallback)")
        dt = logp_t.dtype
        if dt not in (torch.float32, torch.float64):
            raise ValueError("log-probs must be float32 or float64")
        lt = logp_t.detach().contiguous()
        lb = logp_b.detach()
        if lb.dtype != dt:
            lb = lb.to(dt)
        lb = lb.contiguous()
        n = lt.numel()
        if n != handle.n or handle.k1 != k + 1:
            raise ValueError(f"log-probs ({n}) do not match the search result ({handle.n} rows, k1={handle.k1})")
        dev = lt.device
        need_grad = any(ctx.needs_input_grad[:2])
        lib = _lib.load()
        suffix = "f32" if dt == torch.float32 else "f64"
        with _on(dev):
            c, gW, ws = handle.scratch(dt, n_traj)
            scal = torch.empty(8, dtype=torch.float64, device=dev)
            if need_grad:
                rowptr, col = handle.csr()
                unit = torch.empty(n, dtype=dt, device=dev)
                fn = getattr(lib, "mepol_iw_entropy_fwd_bwd_" + suffix)
                rc = fn(lt.data_ptr(), lb.data_ptr(), traj_off.data_ptr(), n_traj, n, handle.idx.data_ptr(),
                        handle.k1, handle.rk(dt).data_ptr(), int(ns), float(log_G), float(B), float(eps),
                        rowptr.data_ptr(), col.data_ptr(), c.data_ptr(), gW.data_ptr(), scal.data_ptr(),
                        unit.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
                _lib.check(rc, "mepol_iw_entropy_fwd_bwd_" + suffix)
                ctx.unit = unit
            else:
                fn = getattr(lib, "mepol_iw_entropy_fwd_" + suffix)
                rc = fn(lt.data_ptr(), lb.data_ptr(), traj_off.data_ptr(), n_traj, n, handle.idx.data_ptr(),
                        handle.k1, handle.rk(dt).data_ptr(), int(ns), float(log_G), float(B), float(eps),
                        c.data_ptr(), None, scal.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
                _lib.check(rc, "mepol_iw_entropy_fwd_" + suffix)
                ctx.unit = None
        # views of the per-call scalar block (no copy kernels); the block is owned by these outputs
        H, kl_raw, n_bad = scal[0], scal[1], scal[5]
        ctx.mark_non_differentiable(kl_raw, n_bad)
        return H, kl_raw, n_bad

    @staticmethod
    def backward(ctx, grad_H, _gkl, _gbad):
        if ctx.unit is None:
            raise RuntimeError("mepol_b200: the objective was evaluated without gradient tracking")
        out = ctx.unit * grad_H          # dH/d logp_target (0-dim dH does not change the storage dtype)
        g_t = out if ctx.needs_input_grad[0] else None
        g_b = -out if ctx.needs_input_grad[1] else None
        return g_t, g_b, None, None, None, None, None, None, None, None


def iw_entropy(logp_t, logp_b, handle, traj_off, n_traj, k, ns, G, B, eps):
    """Device-level objective: returns (H, KL_raw, n_bad_rows) as 0-dim CUDA float64 tensors."""
    return IWEntropyFunction.apply(logp_t, logp_b, handle, traj_off, int(n_traj), int(k), int(ns),
                                   math.lgamma(ns / 2.0 + 1.0) if G is None else math.log(G), float(B), float(eps))


# ------------------------------------------------------------------------------- policy mirror
def _policy_layers(policy):
    """(list of hidden (Linear, activation) pairs, mean Linear, log_std) of a GaussianPolicy-shaped
    module (src/policy.py:17-29): `net` = Sequential(Linear, act, Linear, act, ...), `mean`, `log_std`."""
    mods = list(policy.net)
    return mods, policy.mean, policy.log_std


def mirror_params(policy, device, dtype, requires_grad):
    out = []
    for p in policy.parameters():
        t = p.detach().to(device=device, dtype=dtype)
        out.append(t.requires_grad_(requires_grad))
    return out


class GaussLogPFunction(torch.autograd.Function):
    """log p(a | s) of the diagonal-Gaussian head from the MLP mean: the elementwise tail of
    GaussianPolicy.get_log_p (policy.py:43-51) as ONE kernel forward and one backward
    (mepol_gauss_logp_{fwd,bwd}_{f32,f64}) instead of ~20 ATen elementwise launches on [n, A] temporaries."""

    @staticmethod
    def forward(ctx, mean, actions, log_std):
        if not mean.is_cuda:
            raise _lib.MepolError("GaussLogPFunction needs CUDA tensors (there is no CPU fallback)")
        dt = mean.dtype
        if dt not in (torch.float32, torch.float64):
            raise ValueError("the policy head must be float32 or float64")
        mean_c, act_c, ls_c = mean.detach().contiguous(), actions.detach().to(dt).contiguous(), log_std.detach().to(dt).contiguous()
        n, A = mean_c.shape
        out = torch.empty(n, dtype=dt, device=mean.device)
        suffix = "f32" if dt == torch.float32 else "f64"
        with _on(mean.device):
            rc = getattr(_lib.load(), "mepol_gauss_logp_fwd_" + suffix)(
                mean_c.data_ptr(), act_c.data_ptr(), ls_c.data_ptr(), n, A, out.data_ptr(), _stream_ptr(mean.device))
        _lib.check(rc, "mepol_gauss_logp_fwd_" + suffix)
        ctx.save_for_backward(mean_c, act_c, ls_c)
        ctx.suffix = suffix
        return out

    @staticmethod
    def backward(ctx, g_logp):
        mean_c, act_c, ls_c = ctx.saved_tensors
        n, A = mean_c.shape
        dev, dt = mean_c.device, mean_c.dtype
        g = g_logp.detach().to(dt).contiguous()
        g_mean = torch.empty_like(mean_c)
        g_ls = torch.empty(A, dtype=dt, device=dev)
        ws = torch.empty(_lib.gauss_logp_workspace_bytes(n, A), dtype=torch.uint8, device=dev)
        with _on(dev):
            rc = getattr(_lib.load(), "mepol_gauss_logp_bwd_" + ctx.suffix)(
                mean_c.data_ptr(), act_c.data_ptr(), ls_c.data_ptr(), g.data_ptr(), n, A, g_mean.data_ptr(),
                g_ls.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        _lib.check(rc, "mepol_gauss_logp_bwd_" + ctx.suffix)
        return g_mean, None, g_ls


def gaussian_log_p(policy, params, states, actions):
    """policy.py:43-61 on device tensors with mirrored parameters `params` (ordered like
    policy.parameters(): log_std first, then net.*, then mean.*).  The MLP runs in torch (cuBLAS); on CUDA
    tensors the log-prob tail is the hand-written kernel pair behind GaussLogPFunction.  (CPU tensors --
    only the host-logic test of the parameter ordering uses them -- take the literal expression.)"""
    names = [n for n, _ in policy.named_parameters()]
    table = dict(zip(names, params))
    h = states
    for i, m in enumerate(policy.net):
        if isinstance(m, torch.nn.Linear):
            h = torch.nn.functional.linear(h, table[f"net.{i}.weight"], table[f"net.{i}.bias"])
        else:
            h = m(h)
    mean = torch.nn.functional.linear(h, table["mean.weight"], table["mean.bias"])
    log_std = table["log_std"]
    if mean.is_cuda:
        return GaussLogPFunction.apply(mean, actions, log_std)
    return torch.sum(-0.5 * (_LOG_2PI + 2.0 * log_std
                             + (actions - mean) ** 2 / (torch.exp(log_std) + POLICY_EPS) ** 2), dim=1)


# ------------------------------------------------------------------------------- batch cache
class _Batch:
    def __init__(self, states, actions, real_traj_lengths, num_traj, device, dtype):
        lens = torch.as_tensor(real_traj_lengths).reshape(-1).to(torch.int64).cpu()
        if lens.numel() != num_traj:
            raise ValueError("num_traj does not match real_traj_lengths")
        self.lens = lens.tolist()
        s = torch.as_tensor(states)
        a = torch.as_tensor(actions)
        L = a.shape[1]
        if all(l == L for l in self.lens):
            s_flat = s[:, :L].reshape(-1, s.shape[-1])
            a_flat = a.reshape(-1, a.shape[-1])
        else:
            s_flat = torch.cat([s[t, :l] for t, l in enumerate(self.lens)], 0)
            a_flat = torch.cat([a[t, :l] for t, l in enumerate(self.lens)], 0)
        self.states = s_flat.to(device=device, dtype=dtype).contiguous()
        self.actions = a_flat.to(device=device, dtype=dtype).contiguous()
        self.traj_off, self.n = traj_offsets(lens, device)
        self.n_traj = int(num_traj)
        self.action_dim = int(a.shape[-1])
        self.behav_cache = None     # (param fingerprint, logp_b)
        self.pinned = {}            # pinned host mirrors for async device->host copies


_batches = {}
_handles = {}


_CACHE_CAP = 8


def _cache_prune(table, room_for=1):
    """Drop entries whose key object died, then the oldest ones, so that `room_for` more fit under the
    cap.  Entries hold device memory (indices, reverse CSR, workspaces): the tables must stay bounded
    even when every lookup is a hit on the newest entry (the drop-in attaches one handle per epoch)."""
    for k_ in [k_ for k_, v in table.items() if v[0]() is None]:
        del table[k_]
    while len(table) > _CACHE_CAP - room_for:
        table.pop(next(iter(table)))


def _cache_get(table, key_obj, extra, build):
    key = (id(key_obj), extra)
    hit = table.get(key)
    if hit is not None and hit[0]() is key_obj:
        return hit[1]
    _cache_prune(table)
    val = build()
    table[key] = (weakref.ref(key_obj), val)
    return val


def _get_batch(states, actions, real_traj_lengths, num_traj):
    dev, dt = _device(), _config["compute_dtype"]
    return _cache_get(_batches, states, (id(actions), dt, str(dev), getattr(states, "_version", 0)),
                      lambda: _Batch(states, actions, real_traj_lengths, num_traj, dev, dt))


def attach_handle(indices_cpu, handle):
    """Remember the device-side search result that produced a CPU `indices` tensor (so the objective
    does not have to re-upload it)."""
    _cache_prune(_handles)
    _handles[(id(indices_cpu), handle.k1)] = (weakref.ref(indices_cpu), handle)


def _get_handle(distances, indices, k):
    dev = _device()

    def build():
        idx = torch.as_tensor(indices).to(torch.int32).to(dev)
        rk = torch.as_tensor(distances)[:, k].to(torch.float64).to(dev)
        return SearchHandle(idx, rk, k + 1)

    return _cache_get(_handles, indices, k + 1, build)


def _fingerprint(policy):
    return tuple((p.data_ptr(), p._version) for p in policy.parameters())


_BURN_FAST = None


def _burn_fast_ok():
    """torch's CPU normal_ on a contiguous float64 tensor of n >= 16 elements draws n 64-bit words from
    the generator (+16 when n % 16 != 0: the tail block is redrawn) and nothing else, so the generator
    can be advanced by the same amount without the Box-Muller arithmetic (int64 random_() draws one
    64-bit word per element: 2.5x cheaper).  That is an implementation detail of the installed torch:
    it is CHECKED here once, on a private generator, and the literal randn calls are used if it fails."""
    global _BURN_FAST
    if _BURN_FAST is None:
        ok = True
        for n in (4000, 37 * 3, 16):
            g1, g2 = torch.Generator(), torch.Generator()
            g1.manual_seed(20260101)
            g2.manual_seed(20260101)
            torch.randn(n, dtype=torch.float64, generator=g1)
            torch.empty(n + (16 if n % 16 else 0), dtype=torch.int64).random_(generator=g2)
            ok = ok and torch.equal(g1.get_state(), g2.get_state())
        _BURN_FAST = ok
    return _BURN_FAST


def _burn_rng(batch, calls=2):
    """Q7: every reference get_log_p call draws randn(len, A) from the global CPU generator
    (policy.py:59), target first then behavioural, trajectory by trajectory (mepol.py:127-128)."""
    if not _config["burn_rng"]:
        return
    sizes = [l * batch.action_dim for l in batch.lens]
    if min(sizes) >= 16 and _burn_fast_ok():
        total = calls * sum(n + (16 if n % 16 else 0) for n in sizes)
        buf = batch.pinned.get("burn")
        if buf is None or buf.numel() != total:
            buf = torch.empty(total, dtype=torch.int64)
            batch.pinned["burn"] = buf
        buf.random_()
        return
    for l in batch.lens:
        for _ in range(calls):
            torch.randn((l, batch.action_dim), dtype=torch.float64)


def _behav_logp(behavioral_policy, batch):
    fp = _fingerprint(behavioral_policy)
    if batch.behav_cache is not None and batch.behav_cache[0] == fp:
        return batch.behav_cache[1]
    with torch.no_grad():
        params = mirror_params(behavioral_policy, batch.states.device, batch.states.dtype, False)
        lb = gaussian_log_p(behavioral_policy, params, batch.states, batch.actions)
    batch.behav_cache = (fp, lb)
    return lb


def _pinned(batch, name, like):
    """Per-batch pinned host mirrors (allocated once): async device->host copies need them."""
    buf = batch.pinned.get(name)
    if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
        buf = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
        batch.pinned[name] = buf
    return buf


def _device_step(target_policy, dev_params, lb, batch, handle, k, ns, G, B, eps, want_grad):
    """Forward (+ backward for dH = 1) of the objective on device tensors: (scalars [3], grads|None)."""
    with torch.set_grad_enabled(want_grad):
        lt = gaussian_log_p(target_policy, dev_params, batch.states, batch.actions)
        H, kl_raw, n_bad = iw_entropy(lt, lb, handle, batch.traj_off, batch.n_traj, k, ns, G, B, eps)
    grads = torch.autograd.grad(H, dev_params, allow_unused=True) if want_grad else None
    return torch.stack([H.detach(), kl_raw, n_bad]), grads


class _StaticHandle:
    """Search result living in buffers owned by a captured graph (same interface as SearchHandle)."""

    def __init__(self, handle, dt, want_grad):
        self.idx, self.k1, self.n = torch.empty_like(handle.idx), handle.k1, handle.n
        self._rk = torch.empty_like(handle.rk(dt))
        self._csr = None
        self._scr = None
        if want_grad:          # the reverse CSR only serves the backward: KL / no_grad graphs never build it
            rowptr, col = handle.csr()
            self._csr = (torch.empty_like(rowptr), torch.empty_like(col))

    def load(self, handle, dt):
        self.idx.copy_(handle.idx)
        self._rk.copy_(handle.rk(dt))
        if self._csr is not None:
            rowptr, col = handle.csr()
            self._csr[0].copy_(rowptr)
            self._csr[1].copy_(col)

    def rk(self, dtype):
        return self._rk

    def csr(self):
        return self._csr

    def scratch(self, dtype, n_traj):
        if self._scr is None:        # first (warm-up) run, outside the capture: the buffers persist
            dev = self.idx.device
            self._scr = (torch.empty(self.n, dtype=dtype, device=dev), torch.empty(self.n, dtype=dtype, device=dev),
                         torch.empty(_lib.iw_workspace_bytes(self.n, n_traj), dtype=torch.uint8, device=dev))
        return self._scr


class _StaticBatch:
    def __init__(self, batch, lb):
        self.states, self.actions = torch.empty_like(batch.states), torch.empty_like(batch.actions)
        self.traj_off, self.lb = torch.empty_like(batch.traj_off), torch.empty_like(lb)
        self.n_traj = batch.n_traj

    def load(self, batch, lb):
        self.states.copy_(batch.states)
        self.actions.copy_(batch.actions)
        self.traj_off.copy_(batch.traj_off)
        self.lb.copy_(lb)


class _StepGraph:
    """The device side of one objective step as a replayable CUDA graph: pinned parameter mirrors ->
    device parameters -> MLP -> log-prob -> IW forward (-> IW backward -> MLP backward) -> pinned
    results.  It is captured ONCE per problem shape: states, actions, behavioural log-probs and the
    neighbour lists live in buffers the graph owns, and a new epoch's batch / search result is
    copied into them (a few device-to-device copies) the first time it is seen."""

    def __init__(self, target_policy, cpu_params, lb, batch, handle, k, ns, G, B, eps, want_grad):
        dev, dt = batch.states.device, batch.states.dtype
        self.dt = dt
        self.policy = target_policy
        self.sb, self.sh = _StaticBatch(batch, lb), _StaticHandle(handle, dt, want_grad)
        self.src = None
        self._load(batch, lb, handle)
        self.host_params = [torch.empty(p.shape, dtype=dt, pin_memory=True) for p in cpu_params]
        self.dev_params = [torch.empty(p.shape, dtype=dt, device=dev).requires_grad_(want_grad) for p in cpu_params]
        self.host_out = torch.empty(3, dtype=torch.float64, pin_memory=True)
        self.host_grads = [torch.empty(p.shape, dtype=dt, pin_memory=True) for p in cpu_params] if want_grad else None
        self.want_grad = want_grad
        for hp, p in zip(self.host_params, cpu_params):
            hp.copy_(p.detach())

        def body():
            with torch.no_grad():
                for hp, dp in zip(self.host_params, self.dev_params):
                    dp.copy_(hp, non_blocking=True)
            out, grads = _device_step(target_policy, self.dev_params, self.sb.lb, self.sb, self.sh, k, ns, G, B, eps,
                                      want_grad)
            self.host_out.copy_(out, non_blocking=True)
            self.grad_mask = None if grads is None else [g is not None for g in grads]
            if grads is not None:
                for hg, g in zip(self.host_grads, grads):
                    if g is not None:
                        hg.copy_(g, non_blocking=True)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):          # warm-up outside capture (cuBLAS handles, allocator)
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()

    def _load(self, batch, lb, handle):
        # strong references keep the sources alive, so identity is a valid "already loaded" test
        if self.src is None or self.src[0] is not batch or self.src[1] is not lb:
            self.sb.load(batch, lb)
        if self.src is None or self.src[2] is not handle:
            self.sh.load(handle, self.dt)
        self.src = (batch, lb, handle)

    def run(self, cpu_params, batch, lb, handle):
        self._load(batch, lb, handle)
        for hp, p in zip(self.host_params, cpu_params):
            hp.copy_(p.detach())
        self.graph.replay()
        _burn_rng(batch)                               # host work while the device runs the step
        torch.cuda.current_stream(batch.states.device).synchronize()
        out = self.host_out.clone()
        grads = None
        if self.want_grad:
            grads = [torch.empty(p.shape, dtype=p.dtype).copy_(hg) if m else None
                     for hg, m, p in zip(self.host_grads, self.grad_mask, cpu_params)]
        return out, grads


_graph_failures = 0          # capture is an optimisation: after a few failures stop trying (per process)
_graph_bad_keys = set()
_graph_seen = {}             # key -> sightings: a shape is captured the SECOND time it shows up
_graphs = {}


def _step_graph(target_policy, cpu_params, lb, batch, handle, k, ns, G, B, eps, want_grad):
    """Cached _StepGraph for this problem shape and these objective constants, or None (disabled /
    a capture failed once: the eager path is used from then on)."""
    global _graph_failures
    if not _config["cuda_graphs"] or _graph_failures >= 3:
        return None
    key = (id(target_policy), str(batch.states.device), batch.states.dtype, tuple(batch.states.shape),
           tuple(batch.actions.shape), batch.n_traj, handle.n, handle.k1, bool(want_grad), int(k), int(ns), G,
           float(B), float(eps), tuple(tuple(p.shape) for p in cpu_params))
    if key in _graph_bad_keys:
        return None
    sg = _graphs.pop(key, None)
    if sg is not None and sg.policy is not target_policy:      # id() reuse after the policy was freed
        sg = None
    if sg is None:
        # A capture costs two warm-up runs plus the capture itself.  Shapes that show up once (the
        # full-entropy evaluation, every epoch of an environment with early termination) run eagerly;
        # a shape is captured the second time it is seen.
        seen = _graph_seen.get(key, 0) + 1
        if len(_graph_seen) > 64:
            _graph_seen.clear()
        _graph_seen[key] = seen
        if seen < 2:
            return None
        try:
            sg = _StepGraph(target_policy, cpu_params, lb, batch, handle, k, ns, G, B, eps, want_grad)
        except Exception as e:            # never let a failed capture break a step
            _graph_failures += 1
            _graph_bad_keys.add(key)
            import warnings
            warnings.warn(f"mepol_b200: CUDA-graph capture of the objective step failed ({e!r}); using eager launches")
            torch.cuda.synchronize()
            return None
        while len(_graphs) >= 8:
            _graphs.pop(next(iter(_graphs)))
    _graphs[key] = sg              # (re-)inserted last: eviction is least-recently-used
    return sg


class _EntropyBridge(torch.autograd.Function):
    """CPU parameters -> (entropy, kl_raw, n_bad) CPU scalars, everything in between on the GPU.

    The whole device side of a step -- forward, and when a gradient will be wanted also the backward
    for dH = 1 plus the device->host copies -- is ENQUEUED first (as one CUDA-graph replay when
    graphs are enabled); the Q7 random-number burn then runs on the host while the GPU works, and one
    stream synchronisation ends the call.  backward() only scales the stored gradients by the
    incoming dH (the loss is -H)."""

    @staticmethod
    def forward(ctx, pack, *cpu_params):
        target_policy, lb, batch, handle, k, ns, G, B, eps, grad_mode = pack
        dev, dt = batch.states.device, batch.states.dtype
        want_grad = grad_mode and any(ctx.needs_input_grad[1:])
        with torch.cuda.device(dev):
            sg = _step_graph(target_policy, cpu_params, lb, batch, handle, k, ns, G, B, eps, want_grad)
            if sg is not None:
                out, grads = sg.run(cpu_params, batch, lb, handle)
            else:
                gp = [p.detach().to(device=dev, dtype=dt).requires_grad_(want_grad) for p in cpu_params]
                scal, dgrads = _device_step(target_policy, gp, lb, batch, handle, k, ns, G, B, eps, want_grad)
                out_pin = _pinned(batch, "scalars", scal)
                out_pin.copy_(scal, non_blocking=True)
                pins = None
                if want_grad:
                    pins = []
                    for i, g in enumerate(dgrads):
                        pins.append(None if g is None else _pinned(batch, ("grad", i), g))
                        if g is not None:
                            pins[-1].copy_(g, non_blocking=True)
                _burn_rng(batch)                           # host work while the device runs the step
                torch.cuda.current_stream(dev).synchronize()
                out = out_pin.clone()
                # fresh pageable tensors: the pinned mirrors are reused by the next call on this batch
                grads = None if pins is None else [None if g is None else torch.empty(g.shape, dtype=p.dtype).copy_(g)
                                                   for g, p in zip(pins, cpu_params)]
        ent, klv, bad = out[0].clone(), out[1].clone(), out[2].clone()
        ctx.cpu_grads = grads
        ctx.mark_non_differentiable(klv, bad)
        return ent, klv, bad

    @staticmethod
    def backward(ctx, g_ent, _g1, _g2):
        if ctx.cpu_grads is None:
            raise RuntimeError("mepol_b200: entropy was evaluated without gradient tracking")
        scale = g_ent.to(torch.float64)
        outs = [None if g is None else g * scale.to(g.dtype) for g in ctx.cpu_grads]
        return (None, *outs)


def _evaluate(behavioral_policy, target_policy, states, actions, num_traj, real_traj_lengths,
              distances, indices, k, G, B, ns, eps):
    batch = _get_batch(states, actions, real_traj_lengths, num_traj)
    handle = _get_handle(distances, indices, k)
    lb = _behav_logp(behavioral_policy, batch)
    pack = (target_policy, lb, batch, handle, k, ns, G, B, eps, torch.is_grad_enabled())
    return _EntropyBridge.apply(pack, *list(target_policy.parameters()))


# ------------------------------------------------------------------------------- reference surface
def compute_importance_weights(behavioral_policy, target_policy, states, actions, num_traj, real_traj_lengths):
    """mepol.py:114-139.  Returned for completeness (the objective kernels never materialise w on the
    host): CPU float64 tensor [N] of self-normalised importance weights, no autograd."""
    batch = _get_batch(states, actions, real_traj_lengths, num_traj)
    _burn_rng(batch)
    lb = _behav_logp(behavioral_policy, batch)
    with torch.no_grad():
        params = mirror_params(target_policy, batch.states.device, batch.states.dtype, False)
        lt = gaussian_log_p(target_policy, params, batch.states, batch.actions)
        n = batch.n
        # run the forward kernels for their stabilised log-weights c, M = max c and S = sum exp(c - M);
        # the neighbour inputs are dummies (identity index, unit radius)
        idx = torch.arange(n, dtype=torch.int32, device=lt.device).reshape(n, 1).repeat(1, 2).contiguous()
        rk = torch.ones(n, dtype=lt.dtype, device=lt.device)
        c = torch.empty(n, dtype=lt.dtype, device=lt.device)
        scal = torch.empty(8, dtype=torch.float64, device=lt.device)
        ws = torch.empty(_lib.iw_workspace_bytes(n, batch.n_traj), dtype=torch.uint8, device=lt.device)
        suffix = "f32" if lt.dtype == torch.float32 else "f64"
        with torch.cuda.device(lt.device):
            rc = getattr(_lib.load(), "mepol_iw_entropy_fwd_" + suffix)(
                lt.data_ptr(), lb.to(lt.dtype).contiguous().data_ptr(), batch.traj_off.data_ptr(), batch.n_traj, n,
                idx.data_ptr(), 2, rk.data_ptr(), 1, 0.0, 0.0, 0.0, c.data_ptr(), None,
                scal.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(lt.device))
        _lib.check(rc, "mepol_iw_entropy_fwd_" + suffix)
        w = torch.exp(c.to(torch.float64) - scal[3]) / scal[4]
    return w.cpu()


def compute_entropy(behavioral_policy, target_policy, states, actions,
                    num_traj, real_traj_lengths, distances, indices, k, G, B, ns, eps):
    """mepol.py:142-154 -> 0-dim CPU float64 tensor (differentiable w.r.t. target_policy)."""
    ent, _, _ = _evaluate(behavioral_policy, target_policy, states, actions, num_traj, real_traj_lengths,
                          distances, indices, k, G, B, ns, eps)
    return ent


def compute_kl(behavioral_policy, target_policy, states, actions,
               num_traj, real_traj_lengths, distances, indices, k, eps):
    """mepol.py:157-174 -> (kl clamped at 0, numeric_error) as CPU tensors."""
    with torch.no_grad():
        _, kl_raw, _ = _evaluate(behavioral_policy, target_policy, states, actions, num_traj,
                                 real_traj_lengths, distances, indices, k, None, 0.0, 1, eps)
    numeric_error = torch.isinf(kl_raw) or torch.isnan(kl_raw)
    kl = torch.max(torch.tensor(0.0, dtype=kl_raw.dtype), kl_raw)
    return kl, numeric_error


def policy_update(optimizer, behavioral_policy, target_policy, states, actions,
                  num_traj, traj_len, distances, indices, k, G, B, ns, eps):
    """mepol.py:268-281 (the argument called traj_len receives real_traj_lengths, :429-431)."""
    optimizer.zero_grad()
    loss = - compute_entropy(behavioral_policy, target_policy, states, actions,
                             num_traj, traj_len, distances, indices, k, G, B, ns, eps)
    numeric_error = torch.isinf(loss) or torch.isnan(loss)
    loss.backward()
    optimizer.step()
    return loss, numeric_error
