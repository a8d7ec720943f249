timeout 100 python -m pytest tests/test_gpu_celda_g.py tests/test_gpu_module_edges.py -x -q 2>&1 | tail -2
for c in 50000 500000; do
timeout 300 python scripts/bench_g.py --cells $c --steps 2 --warmup 0 --flip 0 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:60], d['sweep_y_ms'], d['movers_per_sweep'], d['tasks_per_sweep'], d['cta0_mclk_per_sweep'])"
done
