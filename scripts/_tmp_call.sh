timeout 200 python -m pytest tests/test_gpu_celda_c_gibbs.py -x -q 2>&1 | tail -2
timeout 200 python scripts/bench_c_gibbs.py --steps 3 --warmup 1 2>&1 | tail -c 900
