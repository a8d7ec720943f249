set -x
timeout 300 python scripts/bench_g.py --steps 2 --warmup 0 --flip 0 > gpurun_out/g_c4_final.json 2>&1; tail -c 800 gpurun_out/g_c4_final.json
timeout 100 python scripts/bench_g.py --cells 50000 --steps 2 --warmup 0 --flip 0 > gpurun_out/g_50k_final.json 2>&1; tail -c 600 gpurun_out/g_50k_final.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep_yg -c 1 -o gpurun_out/r2_yg_final -f python scripts/bench_g.py --cells 8000 --genes 4000 --steps 1 --warmup 0 --flip 0 > gpurun_out/ncu_yg.log 2>&1; tail -2 gpurun_out/ncu_yg.log
timeout 300 ncu --set full --clock-control none -k regex:'k_colsum_twin_v5|k_rowsum_csc_v4' -c 2 -o gpurun_out/r2_agg_final -f python bench.py --steps 1 --warmup 0 --no-regimes --no-legs --no-cpu-baseline > gpurun_out/ncu_agg.log 2>&1; tail -2 gpurun_out/ncu_agg.log
timeout 300 ncu --set full --clock-control none -k regex:k_em_probs_tma -c 1 -o gpurun_out/r2_em_final -f python scripts/bench_em.py --cells 200000 --steps 1 --warmup 0 > gpurun_out/ncu_em.log 2>&1; tail -2 gpurun_out/ncu_em.log
timeout 300 ncu --set full --clock-control none -k regex:k_sweep_zc -c 1 -o gpurun_out/r2_zc_final -f python scripts/bench_c_gibbs.py --cells 4000 --genes 4000 --steps 1 --warmup 0 > gpurun_out/ncu_zc.log 2>&1; tail -2 gpurun_out/ncu_zc.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2_final.csv python bench.py --steps 2 --warmup 3 --no-regimes --no-legs --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
ls -la gpurun_out/*.ncu-rep
