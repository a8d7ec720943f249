set -x
timeout 300 python -m pytest tests/test_gpu_celda_g.py tests/test_gpu_module_edges.py -x -q 2>&1 | tail -4
timeout 200 python scripts/bench_g.py --cells 50000 --steps 1 --warmup 0 --flip 0 > gpurun_out/g_50k_v4.json 2>&1; tail -c 900 gpurun_out/g_50k_v4.json
timeout 300 python scripts/bench_g.py --steps 1 --warmup 0 --flip 0 > gpurun_out/g_c4_v4.json 2>&1; tail -c 900 gpurun_out/g_c4_v4.json
