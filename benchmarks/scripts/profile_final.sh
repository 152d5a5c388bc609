# Evidence for the FINAL round-1 code (B200, one GPU): ncu launch list of the bench command and `--set full` captures
# of the dominant kernel (K2 CTA pair, one B=4096 step) and of one batch-1 search.  Each ncu pass only after the
# plain command exited 0.  Numbers printed under ncu are never bench values.
set -x
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/plain_final.log 2>&1 && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_list_final.log 2>&1; echo rc=$?
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel_2cta -s 16 -c 4 -f -o gpurun_out/prof_final_gemm2 python bench.py --steps 2 --warmup 3 --skip-cpu --skip-b1 > gpurun_out/ncu_gemm2_final.log 2>&1; echo rc=$?
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk_kernel$|pool_compact_kernel|wide_merge_kernel|rescore_kernel|select_kernel|prep_queries_kernel' -s 80 -c 8 -f -o gpurun_out/prof_final_b1 python benchmarks/sweep.py --rows 1000000 --batches 1 --modes f32 --iters 20 > gpurun_out/ncu_b1_final.log 2>&1; echo rc=$?
