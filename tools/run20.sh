set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_intra.py tests/test_frame.py tests/test_itx.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.txt
for cfg in "48 8" "96 4" "24 16" "96 8"; do set -- $cfg; B200_INTRA_FPS=$1 B200_INTRA_GRID=$2 timeout 600 python bench.py --workload 1080p8_intra --steps 6 --warmup 2 > gpurun_out/bench_intrasb_$1_$2.json 2> gpurun_out/bench_intrasb_$1_$2.err; done
timeout 300 python bench.py --workload itx8x8 > gpurun_out/bench_itx8x8.json 2> gpurun_out/bench_itx8x8.err
tail -3 gpurun_out/pytest_gpu.txt
