#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list (+ full capture of one frame).
# usage: tools/gpu_check.sh [tests|bench|ncu|all]   (outputs under gpurun_out/)
set -u
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; which nasm >> gpurun_out/nproc.txt 2>&1 || echo "no nasm" >> gpurun_out/nproc.txt
if [[ $what == all || $what == tests ]]; then
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
  timeout 600 python bench.py --workload itx8x8 > gpurun_out/bench_itx8x8.json 2> gpurun_out/bench_itx8x8.err
  timeout 900 python bench.py --impl reference --workload 4k10_full --steps 2 --warmup 1 > gpurun_out/bench_ref_4k10.json 2> gpurun_out/bench_ref_4k10.err
  timeout 900 python bench.py --workload 4k10_full > gpurun_out/bench_4k10.json 2> gpurun_out/bench_4k10.err
  timeout 900 python bench.py --impl reference --workload 1080p8_intra --steps 2 --warmup 1 > gpurun_out/bench_ref_intra.json 2> gpurun_out/bench_ref_intra.err
  timeout 900 python bench.py --workload 1080p8_intra > gpurun_out/bench_intra.json 2> gpurun_out/bench_intra.err
fi
if [[ $what == all || $what == ncu ]]; then
  # launch list of 3 timed frames after the warm-up (5 warm-up frames x 10 launches are skipped)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 30 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 5 > gpurun_out/ncu_bench.log 2>&1
  # full capture of one whole frame's kernels
  timeout 1200 ncu --set full --clock-control none --import-source on -s 50 -c 10 \
      -f -o gpurun_out/prof_frame python bench.py --steps 3 --warmup 5 > gpurun_out/ncu_full.log 2>&1
fi
echo done > gpurun_out/done.txt
tail -5 gpurun_out/pytest_gpu.txt 2>/dev/null; tail -3 gpurun_out/smoke.txt 2>/dev/null; cat gpurun_out/bench.json 2>/dev/null; true
