#!/bin/bash
# Round-end measurement: GPU test suite + the bench line of every config (launch-site profile next to each).   tools/final_measure.sh
mkdir -p gpurun_out/final
(timeout 500 python -m pytest tests -m gpu -x -q) > gpurun_out/final/pytest.log 2>&1; tail -4 gpurun_out/final/pytest.log
for c in 2 1 3 4 5; do
  python bench.py --config $c --steps 5 --warmup 3 --dump-sites gpurun_out/final/sites_c$c.txt > gpurun_out/final/bench_c$c.json 2> gpurun_out/final/bench_c$c.err
  echo "config $c rc=$?"; cut -c1-260 gpurun_out/final/bench_c$c.json
done
