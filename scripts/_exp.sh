for m in 0 0x55555555 0x77777777 0xffffffff 0x11111111; do
CELDA_CUDA_LGMASK=$m timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$m', d['ms_per_step'], d['kernel_ms_per_step']['sweep_y'], d['sweep_stats_per_step']['sweep_y']['cta0_mclk'], d['logLik_first_last'])"
done
