"""profiles/r02_sass_knn_tc2.txt: tensor-core / bulk-copy / packed-min mnemonic counts of every pruning-kernel
instantiation in the built library, plus the issuer loop and the epilogue loop head of the shipped one
(python scripts/sass_summary.py; needs cuobjdump, no GPU)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mepol_b200", "libmepol_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
lines = ["""# cuobjdump -sass mepol_b200/libmepol_b200.so -- tensor-core / bulk-copy / packed-min instruction counts per kernel (END of round 2 build)
# mnemonics: UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld (PACK16BIT = .pack::16b),
#            UBLKCP = cp.async.bulk (1-D bulk copy; there is no tensor-map UTMALDG: operands are pre-tiled in the UMMA
#            canonical layout by knn_prep.cu), VHMNMX = 3-input min.f16x2, HMNMX2 = 2-input min.f16x2, FMNMX3 = 3-input min.f32,
#            HSET2 = packed fp16 compare (survivor masks), SYNCS = mbarrier ops, UTCATOMSWS = tcgen05.alloc/dealloc/relinquish,
#            USETMAXREG = setmaxnreg, S2R / S2UR = special-register reads (tid, cluster CTA id)
# template arguments: knn_prune_tc_kernel<kDump, K1, kVar> (round 1, fp32 accumulators);
#                     knn_prune_tc2_kernel<kDump, K1, NC, kFuse, kDbg> (round 2, fp16 accumulators; <false,5,256,1,0> is what the bench runs)
"""]
pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)")
keep = ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "VHMNMX", "HMNMX2", "FMNMX3", "HSET2", "SYNCS", "UTCATOMSWS", "USETMAXREG", "S2R", "S2UR")
shipped = None
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    if "knn_prune_tc" not in name:
        continue
    cnt = collections.Counter()
    for l in f.split("\n"):
        m = pat.match(l)
        if m:
            op = m.group(1)
            for k in keep:
                if op == k or op.startswith(k + "."):
                    cnt[op if k == "LDTM" else k] += 1
    lines.append(name)
    lines.append("    " + "  ".join(f"{k}={v}" for k, v in sorted(cnt.items())))
    if "ILb0ELi5ELi256ELi1ELi0E" in name:
        shipped = f
ls = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in shipped.split("\n") if re.match(r"^\s+/\*[0-9a-f]+\*/", l)]
mma = [i for i, l in enumerate(ls) if "UTCHMMA" in l]
lines.append("\n# issuer loop of knn_prune_tc2_kernel<false,5,256,1,0> (two K steps): two uniform-address TRYWAITs in sequence (stage, then accumulator), 2 UTCHMMA + UTCBAR:")
lines += ls[mma[-2] - 24:mma[-2] + 12]
ld = [i for i, l in enumerate(ls) if "LDTM" in l][-1]
lines.append("\n# epilogue loop head: commit wait, the 256-column packed TMEM load, release of the accumulator (per-warp constants pinned: no S2R / S2UR between them), deferred append, head of the packed 3-input min tree:")
lines += ls[ld - 16:ld + 60]
open(os.path.join(ROOT, "profiles", "r02_sass_knn_tc2.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(l for l in lines if "ILb0ELi5ELi256ELi1ELi0E" in l or l.startswith("    ")) [-400:])
