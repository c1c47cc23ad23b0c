set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
timeout 600 python tools/tune_stage.py 8 dav1d_b200/libb200av1.so dav1d_b200/libb200av1_lr4.so dav1d_b200/libb200av1_lr5.so > gpurun_out/tune8.txt 2>&1
for cfg in "96 8" "48 12" "144 5" "96 6"; do set -- $cfg; B200_INTRA_FPS=$1 B200_INTRA_GRID=$2 timeout 600 python bench.py --workload 1080p8_intra --steps 8 --warmup 3 > gpurun_out/bench_intra_$1_$2.json 2> gpurun_out/bench_intra_$1_$2.err; done
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/tune8.txt
