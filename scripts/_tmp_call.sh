set -x
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_r_shim.py tests/test_gpu_philox.py -m gpu -x -q 2>&1 | tail -5
timeout 80 python bench.py --steps 20 --warmup 3 --no-regimes --no-legs --no-cpu-baseline > gpurun_out/bench_e2e_overlap.json 2> gpurun_out/bench_e2e_overlap.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_e2e_overlap.json'))
print(d['value'], d['ms_per_step'], d['e2e'])
P
