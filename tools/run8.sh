set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --workload 4k10_full > gpurun_out/bench_4k10.json 2> gpurun_out/bench_4k10.err
for cfg in "24 24" "48 12" "16 36" "32 16" "64 8"; do set -- $cfg; B200_INTRA_FPS=$1 B200_INTRA_GRID=$2 timeout 600 python bench.py --workload 1080p8_intra --steps 10 --warmup 3 > gpurun_out/bench_intra_$1_$2.json 2> gpurun_out/bench_intra_$1_$2.err; done
tail -3 gpurun_out/pytest_gpu.txt
