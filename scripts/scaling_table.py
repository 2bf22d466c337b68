"""profiles/r02_scaling.md from the committed bench lines and sweeps (python scripts/scaling_table.py)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
rows = [(g, json.load(open(os.path.join(P, f"r02_bench_n{g}.json")))) for g in (1, 2, 4, 8)]
t1, e1 = rows[0][1]["ms_per_step"], rows[0][1]["e2e"]["ms_per_step"]
c1 = rows[0][1]["clocks"]["sm_mhz"]
out = ["# Round 2 (end) — strong scaling at N=1,048,576, d=29, k=4 (B200, NVLink/NVSwitch box)\n",
       "`python bench.py --gpus G --steps 5 --warmup 3` (torchrun for G>1), device-timed step = sharded search + reverse CSR +",
       "IW forward/backward on rank 0; efficiency = t(1) / (G · t(G)). Every G>1 line carries",
       "`parity: {rows_checked: 1048576, mismatch: 0, rk_mismatch: 0, host_arrays_equal_device_result: true}` (rank 0 repeats the",
       "search alone and compares all rows bitwise; the end-to-end arm's host arrays are compared with the device result).\n",
       "| GPUs | step ms | M states/s | efficiency | pruning kernel ms / rank | frac of sustained peak / GPU | e2e step ms | e2e M states/s | e2e efficiency | SM MHz |",
       "|---|---|---|---|---|---|---|---|---|---|"]
for g, r in rows:
    out.append(f"| {g} | {r['ms_per_step']:.2f} | {r['value'] / 1e6:.2f} | {t1 / (g * r['ms_per_step']):.3f} | "
               f"{r['roofline']['kernel_ms']:.2f} | {r['roofline']['frac']:.3f} | {r['e2e']['ms_per_step']:.2f} | "
               f"{r['e2e']['value'] / 1e6:.2f} | {e1 / (g * r['e2e']['ms_per_step']):.3f} | "
               f"{r['clocks']['sm_mhz']:.0f} ({', '.join(r['clocks']['reasons']) or 'no throttle reason'}) |")
r8 = rows[3][1]
ph, c8 = r8["components"]["phases_ms"], r8["components"]
k1, k8, c8m = rows[0][1]["roofline"]["kernel_ms"], r8["roofline"]["kernel_ms"], r8["clocks"]["sm_mhz"]
eq = (k1 * c1 / c8m + (t1 - k1)) / (8 * r8["ms_per_step"])
out += ["",
        f"Per-phase at 8 GPUs (max over ranks, ms): all-gather of the particles {ph['gather_particles_ms']:.2f}, search of the own rows "
        f"{ph['search_own_rows_ms']:.2f} (pruning kernel {ph['prune_kernel_ms']:.2f}), all-gather of idx / r_k {ph['gather_results_ms']:.2f} "
        f"(includes waiting for the slowest rank), rank 0: reverse CSR {c8['rev_csr_ms']:.2f}, objective {c8['iw_fused_call_ms']:.2f} (one C call).",
        f"The single-GPU row runs at the board's software power cap ({c1:.0f} MHz), the 8-GPU rows near {c8m:.0f} MHz: the efficiency "
        f"column is flattered by clocks (with the single-GPU kernel scaled to the 8-GPU clock it would read {eq:.2f}); what does not shrink "
        f"with G is the replicated operand packing, the two collectives and rank 0's objective, ≈1.0 ms against an {k8:.1f} ms kernel.",
        "The 1-GPU and 8-GPU lines were taken back to back on the same box (GPU 0 of the 8); the 2- and 4-GPU lines on other boxes of the pool, "
        "one commit earlier (log-prob H2D not yet under the search in the end-to-end arm). All four predate the last two kernel changes of the "
        "round (no barrier probes in the epilogue, per-warp constants pinned in registers: together -6.5 % at N = 1M; "
        "`r02_bench_n1_final_kernel.json` is the 1-GPU line with them: 67.7 ms, 15.48 M states/s, 0.70 of the peak). The 1-GPU row of the "
        "sweep table below IS the final kernel; the 2/4/8-GPU rows are not.",
        "",
        "Sweep (`r02_sweep_{1,2,4,8}gpu.jsonl`, sampled-row oracle parity on every line; the 1-GPU file also carries a full sklearn "
        "search per N ≤ 262,144): per-GPU fraction of the sustained peak at N = 64k … 4M:",
        "",
        "| GPUs | 65,536 | 131,072 | 262,144 | 524,288 | 1,048,576 | 2,097,152 | 4,194,304 |",
        "|---|---|---|---|---|---|---|---|"]
for g in (1, 2, 4, 8):
    fr = [json.loads(l)["frac_of_sustained_peak"] for l in open(os.path.join(P, f"r02_sweep_{g}gpu.jsonl")) if l.startswith("{")]
    out.append(f"| {g} | " + " | ".join(f"{f:.2f}" for f in fr) + " |")
open(os.path.join(P, "r02_scaling.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[5:12]))
