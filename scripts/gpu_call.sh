set -x
timeout 600 python -m pytest tests/test_gpu_celda_g.py tests/test_gpu_parity.py tests/test_gpu_em.py -x -q 2>&1 | tail -8
timeout 300 python -m pytest tests/test_gpu_nearsize.py -x -q 2>&1 | tail -5
timeout 200 python scripts/bench_g.py --cells 50000 --steps 2 --warmup 1 > gpurun_out/g_50k.json 2>&1; tail -c 1500 gpurun_out/g_50k.json
timeout 400 python scripts/bench_g.py --steps 1 --warmup 1 > gpurun_out/g_c4.json 2>&1; tail -c 1500 gpurun_out/g_c4.json
timeout 200 python scripts/bench_em.py --steps 5 --warmup 2 > gpurun_out/em_x1.json 2>&1; tail -c 900 gpurun_out/em_x1.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-regimes --no-legs --no-cpu-baseline > gpurun_out/bench_agg.json 2> gpurun_out/bench_agg.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_agg.json').read().strip().splitlines()[-1]); print(json.dumps(d['roofline_aggregation'])); print(d['ms_per_step'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_rowsum_csc_v4|k_colsum_twin_v4' -c 2 -o gpurun_out/r2_agg -f python bench.py --steps 1 --warmup 0 --no-regimes --no-legs --no-cpu-baseline > gpurun_out/ncu_agg.log 2>&1; tail -2 gpurun_out/ncu_agg.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_em_probs_tma -c 1 -o gpurun_out/r2_em -f python scripts/bench_em.py --cells 200000 --steps 1 --warmup 0 > gpurun_out/ncu_em.log 2>&1; tail -2 gpurun_out/ncu_em.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sweep_yg -c 1 -o gpurun_out/r2_yg -f python scripts/bench_g.py --cells 20000 --steps 1 --warmup 0 > gpurun_out/ncu_yg.log 2>&1; tail -2 gpurun_out/ncu_yg.log
ls -la gpurun_out/*.ncu-rep
