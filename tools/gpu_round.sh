#!/bin/bash
# One gpurun call of the round: GPU parity tests, smoke, the bench lines (ours + reference arm) and the stream workloads.
# usage: tools/gpu_round.sh [tests] [bench] [stream] [ncu]     (outputs under gpurun_out/)
set -u
mkdir -p gpurun_out
what=" ${*:-tests bench stream} "
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; which nasm >> gpurun_out/nproc.txt 2>&1 || echo "no nasm" >> gpurun_out/nproc.txt
if [[ $what == *" tests "* ]]; then
  timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
fi
if [[ $what == *" stream "* ]]; then
  for wl in stream1080p8 stream1080p8_inter stream4k8_inter stream4k10; do
    timeout 300 python bench.py --workload $wl --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$wl.json 2> gpurun_out/bench_ref_$wl.err
    timeout 300 python bench.py --workload $wl --steps 5 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  done
  B200HOOK_WAVE_SORT=0 timeout 300 python bench.py --workload stream1080p8 --steps 5 > gpurun_out/bench_stream1080p8_nosort.json 2> gpurun_out/bench_stream1080p8_nosort.err
fi
if [[ $what == *" bench "* ]]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
  timeout 600 python bench.py --impl reference --workload 1080p8_intra --steps 2 --warmup 1 > gpurun_out/bench_ref_intra.json 2> gpurun_out/bench_ref_intra.err
  timeout 600 python bench.py --workload 1080p8_intra > gpurun_out/bench_intra.json 2> gpurun_out/bench_intra.err
fi
if [[ $what == *" bench2 "* ]]; then
  timeout 600 python bench.py --impl reference --workload 4k10_full --steps 2 --warmup 1 > gpurun_out/bench_ref_4k10.json 2> gpurun_out/bench_ref_4k10.err
  timeout 600 python bench.py --workload 4k10_full > gpurun_out/bench_4k10.json 2> gpurun_out/bench_4k10.err
  timeout 600 python bench.py --workload itx8x8 > gpurun_out/bench_itx8x8.json 2> gpurun_out/bench_itx8x8.err
fi
if [[ $what == *" ncu "* ]]; then
  # one frame at a time: the launch order of a frame is then the stage order (the default bench interleaves two frames)
  export B200_FRAMES_IN_FLIGHT=1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 30 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 5 > gpurun_out/ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -s 50 -c 10 \
      -f -o gpurun_out/prof_frame python bench.py --steps 3 --warmup 5 > gpurun_out/ncu_full.log 2>&1
fi
echo done > gpurun_out/done.txt
tail -5 gpurun_out/pytest_gpu.txt 2>/dev/null; tail -3 gpurun_out/smoke.txt 2>/dev/null
for f in gpurun_out/bench*.json; do echo "$f: $(head -c 400 $f)"; done
for f in gpurun_out/bench*stream*.err; do echo "== $f"; tail -3 $f; done
true
