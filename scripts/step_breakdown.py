import os, sys, time
import numpy as np, scipy.special, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import entropy as E
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from config_timings import Policy
T_, L, F, A, hidden, ns, k, eps = 20, 500, 29, 8, [400, 300], 7, 4, 0.0
rng = np.random.default_rng(1); torch.manual_seed(0)
states = rng.standard_normal((T_, L + 1, F)).astype(np.float32); actions = (0.6 * rng.standard_normal((T_, L, A))).astype(np.float32)
nxt = np.ascontiguousarray(states[:, 1:, :ns].reshape(-1, ns))
dist, idx, _h = mepol_b200.search_particles(nxt, k)
beh, tgt = Policy(hidden, F, A), Policy(hidden, F, A); tgt.load_state_dict(beh.state_dict())
opt = torch.optim.Adam(tgt.parameters(), lr=1e-5)
G = float(scipy.special.gamma(ns / 2 + 1)); B = float(np.log(k) - scipy.special.digamma(k))
ts, ta = torch.tensor(states, dtype=torch.float64), torch.tensor(actions, dtype=torch.float64); tl = torch.full((T_, 1), L, dtype=torch.int64)
args = (ts, ta, T_, tl, dist, idx, k, G, B, ns, eps)
def S(): torch.cuda.synchronize(); return time.perf_counter()
for _ in range(3): mepol_b200.policy_update(opt, beh, tgt, *args)
from torch.profiler import profile, ProfilerActivity
t0 = S()
for _ in range(10): mepol_b200.policy_update(opt, beh, tgt, *args)
t1 = S(); print("policy_update ms", (t1 - t0) * 100)
t0 = S()
for _ in range(10):
    with torch.no_grad(): mepol_b200.compute_kl(beh, tgt, ts, ta, T_, tl, dist, idx, k, eps)
t1 = S(); print("compute_kl ms", (t1 - t0) * 100)
# pieces
batch = E._get_batch(ts, ta, tl, T_); handle = E._get_handle(dist, idx, k)
t0 = S()
for _ in range(10): E._burn_rng(batch)
t1 = S(); print("burn_rng ms", (t1 - t0) * 100)
t0 = S()
for _ in range(10): gp = [p.detach().to(device="cuda", dtype=torch.float64).requires_grad_(True) for p in tgt.parameters()]
t1 = S(); print("mirror params ms", (t1 - t0) * 100)
t0 = S()
for _ in range(10): lt = E.gaussian_log_p(tgt, gp, batch.states, batch.actions)
t1 = S(); print("logp fwd ms", (t1 - t0) * 100)
lb = E._behav_logp(beh, batch)
t0 = S()
for _ in range(10): H, klr, nb = E.iw_entropy(lt, lb, handle, batch.traj_off, batch.n_traj, k, ns, G, B, eps)
t1 = S(); print("iw fwd ms", (t1 - t0) * 100)
t0 = S()
for _ in range(10): g = torch.autograd.grad(H, gp, retain_graph=True)
t1 = S(); print("autograd bwd ms", (t1 - t0) * 100)
t0 = S()
for _ in range(10): outs = [x.cpu() for x in g]
t1 = S(); print("grads d2h ms", (t1 - t0) * 100)
t0 = S()
for _ in range(10): out = torch.stack([H.detach(), klr, nb]).cpu()
t1 = S(); print("scalars d2h ms", (t1 - t0) * 100)
t0 = S()
for _ in range(10): opt.step()
t1 = S(); print("cpu adam step ms", (t1 - t0) * 100)
