#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list (+ optional full capture).
# usage: tools/gpu_check.sh [tests|bench|ncu|all]   (outputs under gpurun_out/)
set -u
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; which nasm >> gpurun_out/nproc.txt 2>&1 || echo "no nasm" >> gpurun_out/nproc.txt
if [[ $what == all || $what == tests ]]; then
  timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
fi
if [[ $what == all || $what == bench ]]; then
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:itx_add_kernel -s 3 -c 2 \
      -f -o gpurun_out/prof_itx python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1
fi
echo done > gpurun_out/done.txt
tail -5 gpurun_out/pytest_gpu.txt 2>/dev/null; cat gpurun_out/smoke.txt 2>/dev/null | tail -3; cat gpurun_out/bench.json 2>/dev/null
