# Evidence run with the cluster-launch-control scheduler in the CTA-pair K2 kernel (B200, one GPU).
set -x
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; echo rc=$?
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel_2cta -s 16 -c 4 -f -o gpurun_out/prof_r01i_gemm2 python bench.py --steps 2 --warmup 3 --skip-cpu --skip-b1 > gpurun_out/ncu_gemm2_r01i.log 2>&1; echo rc=$?
