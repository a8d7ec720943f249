for f in 3 67; do
for c in 50000 500000; do
CELDA_CUDA_YG_FLAGS=$f timeout 300 python scripts/bench_g.py --cells $c --steps 2 --warmup 0 --flip 0 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('flags $f', d['config']['workload'][:60], d['sweep_y_ms'], d['movers_per_sweep'], d['cta0_mclk_per_sweep'])"
done
done
