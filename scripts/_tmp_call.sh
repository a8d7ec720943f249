set -x
timeout 80 python bench.py --steps 20 --warmup 3 --no-regimes --no-legs --no-cpu-baseline > gpurun_out/bench_philox_leg.json 2> gpurun_out/bench_philox_leg.err
tail -c 400 gpurun_out/bench_philox_leg.err
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_philox_leg.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['philox_mode'])
P
