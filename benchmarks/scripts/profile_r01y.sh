# ncu --set full of every kernel of a batch-1 search with the wide first slab (B200, one GPU)
set -x
timeout -s KILL 300 python benchmarks/sweep.py --rows 1000000 --batches 1 --modes f32 --iters 20 > gpurun_out/plain_r01y.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk_kernel$|pool_compact_kernel|wide_merge_kernel|rescore_kernel|select_kernel|prep_queries_kernel' -s 80 -c 8 -f -o gpurun_out/prof_r01y_b1 python benchmarks/sweep.py --rows 1000000 --batches 1 --modes f32 --iters 20 > gpurun_out/ncu_b1_r01y.log 2>&1; echo rc=$?
