for kb in 200 207 214; do for mc in 31 1; do
CELDA_CUDA_ZSMEM_KB=$kb CELDA_CUDA_MC=$mc timeout 300 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$kb $mc', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['kernel_ms_per_step'].items() if v>0.1}, d['sweep_stats_per_step']['sweep_z']['cta0_mclk'], d['sweep_stats_per_step']['sweep_y']['rounds'])"
done; done
