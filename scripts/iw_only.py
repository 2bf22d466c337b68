"""Runs the IW forward/backward kernels alone (C ABI, pre-allocated buffers) on the benchmark shape, for ncu."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mepol_b200
from mepol_b200 import _lib
from mepol_b200 import entropy as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
L, k, d = 500, 4, 29
rng = np.random.default_rng(0)
lens = [L] * (n // L) + ([n % L] if n % L else [])
dev = torch.device("cuda")
idx = torch.from_numpy(rng.integers(0, n, (n, k + 1)).astype(np.int32)).to(dev); idx[:, 0] = torch.arange(n, device=dev, dtype=torch.int32)
rk = torch.from_numpy(rng.uniform(3, 6, n)).to(dev)
h = E.SearchHandle(idx, rk, k + 1)
off, _ = mepol_b200.traj_offsets(np.asarray(lens), dev)
lt = torch.from_numpy(rng.normal(0, 0.01, n).astype(np.float32)).to(dev); lb = torch.zeros(n, device=dev)
lib = _lib.load(); nt = len(lens)
c = torch.empty(n, device=dev); gw = torch.empty(n, device=dev); gr = torch.empty(n, device=dev)
sc = torch.empty(8, dtype=torch.float64, device=dev); one = torch.ones(1, dtype=torch.float64, device=dev)
ws = torch.empty(_lib.iw_workspace_bytes(n, nt), dtype=torch.uint8, device=dev)
rowptr, col = h.csr(); rk32 = h.rk(torch.float32); sp = torch.cuda.current_stream().cuda_stream
G = math.lgamma(d / 2 + 1)
for _ in range(3):
    _lib.check(lib.mepol_iw_entropy_fwd_f32(lt.data_ptr(), lb.data_ptr(), off.data_ptr(), nt, n, idx.data_ptr(), k + 1, rk32.data_ptr(), d, G, 0.1, 0.0,
                                            c.data_ptr(), gw.data_ptr(), sc.data_ptr(), ws.data_ptr(), ws.numel(), sp), "fwd")
    _lib.check(lib.mepol_iw_entropy_bwd_f32(c.data_ptr(), gw.data_ptr(), sc.data_ptr(), off.data_ptr(), nt, n, rowptr.data_ptr(), col.data_ptr(), one.data_ptr(),
                                            gr.data_ptr(), sp), "bwd")
torch.cuda.synchronize()
print("H", sc[0].item(), "KL", sc[1].item())
if "--prof" in sys.argv:      # per-kernel device times (CUPTI), 20 forward + backward pairs
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(20):
            _lib.check(lib.mepol_iw_entropy_fwd_f32(lt.data_ptr(), lb.data_ptr(), off.data_ptr(), nt, n, idx.data_ptr(), k + 1, rk32.data_ptr(), d, G, 0.1, 0.0,
                                                    c.data_ptr(), gw.data_ptr(), sc.data_ptr(), ws.data_ptr(), ws.numel(), sp), "fwd")
            _lib.check(lib.mepol_iw_entropy_bwd_f32(c.data_ptr(), gw.data_ptr(), sc.data_ptr(), off.data_ptr(), nt, n, rowptr.data_ptr(), col.data_ptr(), one.data_ptr(),
                                                    gr.data_ptr(), sp), "bwd")
        torch.cuda.synchronize()
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
        if e.device_time_total > 0:
            print(f"{e.key[:70]:70s} n={e.count:3d} mean_us={e.device_time_total / e.count:8.2f}")
