#!/bin/bash
# BASELINE.json configs[0..3] (+ the k=4 / 29-dim readings) x seeds {0,1,2} through bench.py, one JSON line each.
#   bash scripts/bench_configs.sh > profiles/r02_bench_configs.jsonl
cd "$(dirname "$0")/.."
for w in grid_world grid_world_k4 mountain_car ant ant_29d humanoid; do
  for s in 0 1 2; do
    extra="--no-cpu-baseline"
    [ "$s" = "0" ] && extra=""
    python bench.py --workload $w --seed $s --steps 20 --warmup 5 $extra 2>/dev/null | grep '^{'
  done
done
