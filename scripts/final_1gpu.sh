#!/bin/bash
# End-of-session 1-GPU measurement pass (run under gpurun from the repo root): GPU tests, the bench line, the ncu
# launch list of the bench, one ncu --set full capture of the pruning kernel and of the IW kernels, the shipped
# configurations.  Everything lands in gpurun_out/ (copied into profiles/ by hand afterwards).
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/f_pytest_gpu.txt 2>&1; tail -2 $O/f_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > $O/f_bench_n1.json 2> $O/f_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/f_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > $O/f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_prune_tc2_kernel --launch-skip 2 --launch-count 1 \
    -f -o $O/f_tc2 python scripts/exp_variants.py 3 1048576 > $O/f_ncu_tc2.log 2>&1
ncu -i $O/f_tc2.ncu-rep --page raw --csv > $O/f_ncu_tc2_raw.csv 2>/dev/null
ncu --set full --clock-control none -k regex:iw_ --launch-skip 8 --launch-count 4 -f -o $O/f_iw python scripts/iw_only.py > $O/f_ncu_iw.log 2>&1
ncu -i $O/f_iw.ncu-rep --page raw --csv > $O/f_ncu_iw_raw.csv 2>/dev/null
bash scripts/bench_configs.sh > $O/f_bench_configs.jsonl
python bench.py --impl reference --steps 2 --warmup 1 > $O/f_bench_ref_n1.json 2> $O/f_bench_ref_n1.err
ls -la $O | grep " f_"
