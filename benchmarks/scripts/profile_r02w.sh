# ncu evidence for the final round-2 code (after the compaction staging, the run-aggregated flush, the group widths) (B200, one GPU).  Each ncu pass only after the plain command exited 0; numbers
# printed under ncu are never bench values.  Summaries: benchmarks/ncu_summarise.py -> profiles/r02w_*.
set -x
P="python benchmarks/profile_cases.py"
timeout -s KILL 600 python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras > gpurun_out/plain_r02w.log 2>&1 && \
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02w.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-extras > gpurun_out/ncu_list_r02w.log 2>&1; echo rc=$?
timeout -s KILL 300 $P step > gpurun_out/plain_step.log 2>&1 && \
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemm_topk_kernel_2cta -s 16 -c 4 -f -o gpurun_out/prof_r02w_gemm2 $P step > gpurun_out/ncu_step.log 2>&1; echo rc=$?
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:pool_compact_kernel|rescore_kernel|select_kernel|prep_queries_kernel' -s 28 -c 7 -f -o gpurun_out/prof_r02w_tail $P step > gpurun_out/ncu_tail.log 2>&1; echo rc=$?
timeout -s KILL 300 $P b1 > gpurun_out/plain_b1.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk_kernel$|pool_compact_kernel|wide_merge_kernel|rescore_kernel|select_kernel|prep_queries_kernel' -s 72 -c 8 -f -o gpurun_out/prof_r02w_b1 $P b1 > gpurun_out/ncu_b1.log 2>&1; echo rc=$?
timeout -s KILL 300 $P tf32 > gpurun_out/plain_tf32.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm_topk_kernel$' -s 16 -c 2 -f -o gpurun_out/prof_r02w_tf32 $P tf32 > gpurun_out/ncu_tf32.log 2>&1; echo rc=$?
timeout -s KILL 300 $P scan > gpurun_out/plain_scan.log 2>&1 && \
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k 'regex:scan_kernel' -s 16 -c 2 -f -o gpurun_out/prof_r02w_scan $P scan > gpurun_out/ncu_scan.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
