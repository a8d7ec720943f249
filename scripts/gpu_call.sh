set -x
timeout 200 python scripts/bench_g.py --cells 50000 --steps 1 --warmup 0 --flip 0 > gpurun_out/g_50k_seg0.json 2>&1; tail -c 1600 gpurun_out/g_50k_seg0.json
