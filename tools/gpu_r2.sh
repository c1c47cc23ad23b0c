#!/bin/bash
# Round-2 gpurun driver. usage: tools/gpu_r2.sh <section> ...   (outputs under gpurun_out/)
set -u
mkdir -p gpurun_out
what=" $* "
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/nproc.txt 2>&1
run_bench() { # name, env..., -- args
  local name=$1; shift
  ( timeout 600 env "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err ) ; echo "bench_$name rc=$?"
}
if [[ $what == *" probe "* ]]; then
  timeout 120 python tools/peer_probe.py 2 > gpurun_out/probe_waitvalue.txt 2>&1; echo "probe rc=$?"; tail -12 gpurun_out/probe_waitvalue.txt
  B200_FLAG_KERNELS=1 timeout 120 python tools/peer_probe.py 2 > gpurun_out/probe_kernels.txt 2>&1; echo "probe(kernels) rc=$?"; tail -12 gpurun_out/probe_kernels.txt
fi
if [[ $what == *" gopprobe "* ]]; then
  IFS=';' read -ra CFGS <<< "${B200_PROBE_CFGS:-2 1 0;2 1 1;2 2 1}"
  for cfg in "${CFGS[@]}"; do
    timeout 90 python tools/gop_probe.py $cfg > "gpurun_out/gop_probe_${cfg// /_}.txt" 2>&1; echo "gop_probe $cfg rc=$?"; grep -E "PARITY|STALL|flags|synced|Error|error" "gpurun_out/gop_probe_${cfg// /_}.txt" | head -8
  done
fi
if [[ $what == *" newtests "* ]]; then
  timeout 600 python -m pytest tests/test_multigpu.py -x -v -m gpu --timeout 150 > gpurun_out/pytest_multigpu.txt 2>&1; echo "multigpu tests rc=$?"; tail -15 gpurun_out/pytest_multigpu.txt
  timeout 600 python -m pytest tests/test_frame.py tests/test_mc.py -x -q -m gpu --timeout 300 > gpurun_out/pytest_new.txt 2>&1; echo "frame/mc tests rc=$?"; tail -5 gpurun_out/pytest_new.txt
fi
if [[ $what == *" tests "* ]]; then
  timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
  tail -5 gpurun_out/pytest_gpu.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
fi
if [[ $what == *" bench1 "* ]]; then
  run_bench n1_default python bench.py --steps 20 --warmup 5
  run_bench n1_2fl_b512 B200_FRAMES_IN_FLIGHT=2 B200_BAND_ROWS=512 python bench.py --steps 20 --warmup 5
  run_bench n1_2fl_b256 B200_FRAMES_IN_FLIGHT=2 B200_BAND_ROWS=256 python bench.py --steps 20 --warmup 5
  run_bench n1_1fl_b128 B200_FRAMES_IN_FLIGHT=1 B200_BAND_ROWS=128 python bench.py --steps 20 --warmup 5
fi
if [[ $what == *" tests2 "* ]]; then
  timeout 900 python -m pytest tests/test_intra.py tests/test_filmgrain.py tests/test_frame.py tests/test_mc.py tests/test_ipred.py -x -q -m gpu --timeout 300 > gpurun_out/pytest_tests2.txt 2>&1; echo "tests2 rc=$?"; tail -4 gpurun_out/pytest_tests2.txt
fi
if [[ $what == *" benchA "* ]]; then
  run_bench n1_default python bench.py --steps 20 --warmup 5
  run_bench n1_fused B200_FUSED=1 python bench.py --steps 20 --warmup 5
  run_bench intra python bench.py --workload 1080p8_intra --steps 10 --warmup 3
  run_bench intra_cta B200_INTRA_CTA=1 python bench.py --workload 1080p8_intra --steps 10 --warmup 3
  run_bench 4k10 python bench.py --workload 4k10_full --steps 20 --warmup 5
fi
if [[ $what == *" benchB "* ]]; then
  run_bench n1_default python bench.py --steps 20 --warmup 5
  run_bench n1_fused B200_FUSED=1 python bench.py --steps 20 --warmup 5
  run_bench n1_mixed python bench.py --workload 4k8_mixed --steps 20 --warmup 5
  run_bench 4k10 python bench.py --workload 4k10_full --steps 20 --warmup 5
fi
if [[ $what == *" benchC "* ]]; then
  run_bench n1_default python bench.py --steps 20 --warmup 5
  run_bench n1_2fl_b576 B200_FRAMES_IN_FLIGHT=2 B200_BAND_ROWS=576 python bench.py --steps 20 --warmup 5
  run_bench n1_2fl_b384 B200_FRAMES_IN_FLIGHT=2 B200_BAND_ROWS=384 python bench.py --steps 20 --warmup 5
  run_bench n1_2fl_b1088 B200_FRAMES_IN_FLIGHT=2 B200_BAND_ROWS=1088 python bench.py --steps 20 --warmup 5
  run_bench n1_1fl_b320 B200_FRAMES_IN_FLIGHT=1 B200_BAND_ROWS=320 python bench.py --steps 20 --warmup 5
  run_bench n1_1fl_b320_nograph B200_GRAPHS=0 B200_FRAMES_IN_FLIGHT=1 B200_BAND_ROWS=320 python bench.py --steps 20 --warmup 5
  run_bench intra_warp B200_INTRA_SB=0 python bench.py --workload 1080p8_intra --steps 10 --warmup 3
  run_bench intra_warp_g32 B200_INTRA_SB=0 B200_INTRA_GRID=32 python bench.py --workload 1080p8_intra --steps 10 --warmup 3
  run_bench n1_mixed python bench.py --workload 4k8_mixed --steps 20 --warmup 5
fi
if [[ $what == *" benchD "* ]]; then
  run_bench n1_default python bench.py --steps 20 --warmup 5
  run_bench n1_1fl_b320 B200_FRAMES_IN_FLIGHT=1 B200_BAND_ROWS=320 python bench.py --steps 20 --warmup 5
  run_bench n1_2fl_b576 B200_FRAMES_IN_FLIGHT=2 B200_BAND_ROWS=576 python bench.py --steps 20 --warmup 5
  run_bench n1_2fl_b320 B200_FRAMES_IN_FLIGHT=2 B200_BAND_ROWS=320 python bench.py --steps 20 --warmup 5
fi
if [[ $what == *" benchI "* ]]; then
  run_bench intra_warp B200_INTRA_SB=0 python bench.py --workload 1080p8_intra --steps 10 --warmup 3
  run_bench intra_warp_g16 B200_INTRA_SB=0 B200_INTRA_GRID=16 python bench.py --workload 1080p8_intra --steps 10 --warmup 3
  run_bench intra_cta B200_INTRA_SB=0 B200_INTRA_CTA=1 python bench.py --workload 1080p8_intra --steps 10 --warmup 3
fi
if [[ $what == *" profmc "* ]]; then
  export B200_SKIP_PARITY=1 B200_MIN_TIMED_S=0.005 B200_NSETS=3 B200_DISTINCT=1
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:mc_pred_kernel -s 3 -c 2 -f -o gpurun_out/prof_mc python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_mc.log 2>&1; echo "ncu mc rc=$?"
  unset B200_SKIP_PARITY B200_MIN_TIMED_S B200_NSETS B200_DISTINCT
fi
if [[ $what == *" benchref "* ]]; then
  run_bench ref python bench.py --impl reference --steps 3 --warmup 1
fi
if [[ $what == *" mgtests "* ]]; then
  timeout 500 python -m pytest tests/test_multigpu.py -x -v -m gpu --timeout 200 > gpurun_out/pytest_multigpu.txt 2>&1; echo "multigpu tests rc=$?"; tail -8 gpurun_out/pytest_multigpu.txt
fi
if [[ $what == *" bench2gpu "* ]]; then
  for wl in 4k8_inter 8k10_full; do
    ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_n2_$wl.json 2> gpurun_out/bench_n2_$wl.err ); echo "n2 $wl rc=$?"
  done
fi
if [[ $what == *" benchNgpu "* ]]; then
  N=${B200_N:-8}
  for wl in 8k10_full 4k8_inter; do
    ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --workload $wl > gpurun_out/bench_n${N}_$wl.json 2> gpurun_out/bench_n${N}_$wl.err ); echo "n$N $wl rc=$?"
  done
fi
if [[ $what == *" bandsweep "* ]]; then
  # band height vs throughput of the dependent stream at N ranks (the chain bound is ~2 band times per frame)
  N=${B200_N:-4}
  for cfg in "4k8_inter 128" "4k8_inter 192" "4k8_inter 320" "8k10_full 128" "8k10_full 192"; do
    set -- $cfg
    ( B200_SKIP_PARITY=1 B200_BAND_ROWS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 16 --warmup 4 --workload $1 > gpurun_out/bench_n${N}_$1_b$2.json 2> gpurun_out/bench_n${N}_$1_b$2.err ); echo "n$N $1 b$2 rc=$?"
  done
fi
if [[ $what == *" final "* ]]; then
  # round-end pass on one GPU: every GPU test, smoke, the bench lines the docs quote, launch list, ncu --set full of every
  # frame-path kernel (8-bit and 10-bit instantiations, film grain, the intra machine, warp / blend)
  timeout 1200 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
  run_bench n1_default python bench.py
  export B200_SKIP_PARITY=1 B200_MIN_TIMED_S=0.005 B200_NSETS=3 B200_DISTINCT=1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_4k8.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 400 ncu --set full --clock-control none -k regex:"mc_|itx_|lf_|cdef_|lr_|coef_" -s 36 -c 12 -f -o gpurun_out/prof_4k8 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_4k8.log 2>&1; echo "ncu 4k8 rc=$?"
  timeout 400 ncu --set full --clock-control none -k regex:"mc_|itx_|lf_|cdef_|lr_|coef_|fg_" -s 42 -c 14 -f -o gpurun_out/prof_4k10 python bench.py --workload 4k10_full --steps 2 --warmup 3 > gpurun_out/ncu_4k10.log 2>&1; echo "ncu 4k10 rc=$?"
  timeout 500 ncu --set full --clock-control none -k regex:"intra_|mc_warp|mc_blend" -s 12 -c 4 -f -o gpurun_out/prof_mixed python bench.py --workload 4k8_mixed --steps 2 --warmup 3 > gpurun_out/ncu_mixed.log 2>&1; echo "ncu mixed rc=$?"
  unset B200_SKIP_PARITY B200_MIN_TIMED_S B200_NSETS B200_DISTINCT
  run_bench n1_mixed python bench.py --workload 4k8_mixed --steps 20 --warmup 5
  run_bench n1_4k10 python bench.py --workload 4k10_full --steps 20 --warmup 5
  run_bench n1_8k10 python bench.py --workload 8k10_full --steps 10 --warmup 3
  run_bench n1_intra_warp B200_INTRA_SB=0 python bench.py --workload 1080p8_intra --steps 10 --warmup 3
  ls -la gpurun_out/*.ncu-rep
fi
if [[ $what == *" final2 "* ]]; then
  # the GPU tests again (new stream tests), ncu --set full of every frame-path kernel dumped to CSV on the box (the reports
  # themselves are tens of MB), the bench lines in full
  timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
  export B200_SKIP_PARITY=1 B200_MIN_TIMED_S=0.005 B200_NSETS=3 B200_DISTINCT=1
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_4k8.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 400 ncu --set full --clock-control none -k regex:"mc_|itx_|lf_|cdef_|lr_|coef_" -s 36 -c 12 -f -o /tmp/prof_4k8 python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_4k8.log 2>&1; echo "ncu 4k8 rc=$?"
  timeout 400 ncu --set full --clock-control none -k regex:"mc_|itx_|lf_|cdef_|lr_|coef_|fg_" -s 42 -c 14 -f -o /tmp/prof_4k10 python bench.py --workload 4k10_full --steps 2 --warmup 3 > gpurun_out/ncu_4k10.log 2>&1; echo "ncu 4k10 rc=$?"
  timeout 500 ncu --set full --clock-control none -k regex:"intra_|mc_warp|mc_blend" -s 12 -c 4 -f -o /tmp/prof_mixed python bench.py --workload 4k8_mixed --steps 2 --warmup 3 > gpurun_out/ncu_mixed.log 2>&1; echo "ncu mixed rc=$?"
  unset B200_SKIP_PARITY B200_MIN_TIMED_S B200_NSETS B200_DISTINCT
  for n in 4k8 4k10 mixed; do ncu -i /tmp/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_$n.csv 2>/dev/null; done
  ls -la /tmp/*.ncu-rep gpurun_out/*.csv
  run_bench n1_default python bench.py
  run_bench n1_mixed python bench.py --workload 4k8_mixed --steps 20 --warmup 5
  run_bench n1_4k10 python bench.py --workload 4k10_full --steps 20 --warmup 5
  run_bench n1_8k10 python bench.py --workload 8k10_full --steps 10 --warmup 3
  du -sh gpurun_out
fi
if [[ $what == *" async "* ]]; then
  # the hooked decoder with asynchronous submission (events between frame jobs, page-locked output pictures): stream tests,
  # then the stream workloads with and without it, and the stock decoder beside them
  timeout 300 python -m pytest tests/test_stream.py -x -q -m gpu --timeout 120 2>&1 | tail -6 > gpurun_out/pytest_stream_gpu.txt; tail -3 gpurun_out/pytest_stream_gpu.txt
  run_bench stream1080p8_inter python bench.py --workload stream1080p8_inter --steps 3 --warmup 1
  run_bench stream1080p8_inter_sync B200HOOK_ASYNC=0 B200HOOK_PINNED_PICS=0 python bench.py --workload stream1080p8_inter --steps 3 --warmup 1
  run_bench stream4k8_inter python bench.py --workload stream4k8_inter --steps 3 --warmup 1
  run_bench ref_stream1080p8_inter python bench.py --impl reference --workload stream1080p8_inter --steps 3 --warmup 1
fi
if [[ $what == *" async4k "* ]]; then
  run_bench stream4k8_inter_sync B200HOOK_ASYNC=0 B200HOOK_PINNED_PICS=0 python bench.py --workload stream4k8_inter --steps 2 --warmup 1
  run_bench ref_stream4k8_inter python bench.py --impl reference --workload stream4k8_inter --steps 2 --warmup 1
fi
echo done > gpurun_out/done.txt
for f in gpurun_out/bench_*.json; do echo "$f: $(head -c 600 $f)"; done
for f in gpurun_out/bench_*.err; do if [ -s $f ]; then echo "== $f"; tail -5 $f; fi; done
true
