set -u
mkdir -p gpurun_out
timeout 900 python tools/tune_stage.py 8 dav1d_b200/libb200av1.so dav1d_b200/libb200av1_mcE.so dav1d_b200/libb200av1_mcF.so dav1d_b200/libb200av1_mcG.so > gpurun_out/tune8.txt 2>&1
for cfg in "24 24" "31 19" "31 12" "16 36"; do set -- $cfg; B200_INTRA_FPS=$1 B200_INTRA_GRID=$2 timeout 600 python bench.py --workload 1080p8_intra --steps 6 --warmup 3 > gpurun_out/bench_intra_$1_$2.json 2> gpurun_out/bench_intra_$1_$2.err; done
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/tune8.txt
