# Final round-1 evidence run (B200, one GPU).  Numbers under ncu are never bench values.
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01g.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_list_r01g.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel_2cta -s 16 -c 4 -f -o gpurun_out/prof_r01g_gemm2 python bench.py --steps 2 --warmup 3 --skip-cpu --skip-b1 > gpurun_out/ncu_gemm2_r01g.log 2>&1; echo rc=$?
ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk_kernel$' -s 8 -c 4 -f -o gpurun_out/prof_r01g_b1 python bench.py --steps 1 --warmup 3 --skip-cpu > gpurun_out/ncu_b1_r01g.log 2>&1; echo rc=$?
ncu --set full --clock-control none -k 'regex:rescore_kernel|select_kernel|pool_compact_kernel|prep_queries_kernel' -s 34 -c 7 -f -o gpurun_out/prof_r01g_misc python bench.py --steps 2 --warmup 3 --skip-cpu --skip-b1 > gpurun_out/ncu_misc_r01g.log 2>&1; echo rc=$?
