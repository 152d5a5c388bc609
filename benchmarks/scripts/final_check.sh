# What the driver runs at round end, in one go: GPU tests, smoke(), the bench line, the reference arm.
set -x
timeout -s KILL 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -s KILL 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
timeout -s KILL 900 python bench.py --impl reference > gpurun_out/ref_final.json 2> gpurun_out/ref_final.err; echo rc=$?
