set -x
timeout 300 python -m pytest tests/test_gpu_celda_g.py -x -q 2>&1 | tail -12
timeout 200 python scripts/bench_g.py --cells 50000 --steps 2 --warmup 0 --flip 0 --scan-mode 1 > gpurun_out/g_50k_skip.json 2>&1; tail -c 500 gpurun_out/g_50k_skip.json
