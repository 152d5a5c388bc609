# Final round-1 evidence run (B200, one GPU): the bench line, then -- each only after the plain run exited 0 --
# the ncu launch list of the same command and `--set full` captures of the kernels.  Numbers printed under ncu
# are never bench values.
set -x
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01j.json 2> gpurun_out/bench_r01j.err; echo rc=$?
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/plain_r01j.log 2>&1 && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01j.csv python bench.py --steps 2 --warmup 3 --skip-cpu > gpurun_out/ncu_list_r01j.log 2>&1; echo rc=$?
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel_2cta -s 16 -c 4 -f -o gpurun_out/prof_r01j_gemm2 python bench.py --steps 2 --warmup 3 --skip-cpu --skip-b1 > gpurun_out/ncu_gemm2_r01j.log 2>&1; echo rc=$?
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk_kernel$' -s 8 -c 4 -f -o gpurun_out/prof_r01j_b1 python bench.py --steps 1 --warmup 3 --skip-cpu > gpurun_out/ncu_b1_r01j.log 2>&1; echo rc=$?
timeout -s KILL 900 ncu --set full --clock-control none -k 'regex:rescore_kernel|select_kernel|pool_compact_kernel|prep_queries_kernel' -s 34 -c 7 -f -o gpurun_out/prof_r01j_misc python bench.py --steps 2 --warmup 3 --skip-cpu --skip-b1 > gpurun_out/ncu_misc_r01j.log 2>&1; echo rc=$?
timeout -s KILL 900 ncu --set full --clock-control none -k 'regex:scan_kernel' -s 8 -c 4 -f -o gpurun_out/prof_r01j_scan python bench.py --steps 1 --warmup 3 --skip-cpu > gpurun_out/ncu_scan_r01j.log 2>&1; echo rc=$?
