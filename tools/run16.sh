set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_intra.py tests/test_itx.py tests/test_frame.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_gpu.txt
timeout 600 python tools/tune_stage.py 8 dav1d_b200/libb200av1.so dav1d_b200/libb200av1_itx8.so dav1d_b200/libb200av1_itx6.so > gpurun_out/tune8.txt 2>&1
for cfg in "144 5" "96 8"; do set -- $cfg; B200_INTRA_FPS=$1 B200_INTRA_GRID=$2 timeout 600 python bench.py --workload 1080p8_intra --steps 8 --warmup 3 > gpurun_out/bench_intra_$1_$2.json 2> gpurun_out/bench_intra_$1_$2.err; done
# ncu: one LR launch, one mc_pred launch (full sets)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lr_frame_kernel|mc_pred_kernel|itx_add_grouped" -s 15 -c 4 -f -o gpurun_out/prof_lr python bench.py --steps 3 --warmup 5 > gpurun_out/ncu_lr.log 2>&1
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/tune8.txt
