set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?
tail -c 600 gpurun_out/bench_final.json
timeout 600 python bench.py --start random --steps 4 --warmup 0 --no-cpu-baseline > gpurun_out/bench_random2.json 2> gpurun_out/bench_random2.err; echo rc=$?
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err; echo rc=$?
cat gpurun_out/bench_ref2.json | head -c 800
