set -x
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_nearsize.py 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_s2_final.json 2> gpurun_out/bench_s2_final.err; echo rc=$?
tail -c 300 gpurun_out/bench_s2_final.err
