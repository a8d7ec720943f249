# The GPU-box sequence behind the round-2 profiles (run from the repo root through gpurun).
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?          # profiles/r2_bench_1gpu_full.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json                 # profiles/r2_bench_reference_arm.json
timeout 300 python scripts/bench_g.py --steps 2 --warmup 0 --flip 0 > gpurun_out/g_c4.json                    # profiles/r2_celdaG_config4.json
timeout 300 python scripts/bench_g.py --steps 2 --warmup 0 --flip 0 --scan-mode 1 > gpurun_out/g_c4_skip.json # profiles/r2_celdaG_config4_scanmode1.json
timeout 200 python scripts/bench_em.py --steps 5 --warmup 2 > gpurun_out/em.json                              # profiles/r2_em_config3_1gpu.json
# ncu (one capture per kernel, after the plain run exited 0); summaries: python scripts/ncu_summary.py <rep> profiles/<name>_summary.csv
timeout 300 ncu --set full --clock-control none -k regex:'k_colsum_twin_v5|k_rowsum_csc_v4' -c 2 -o gpurun_out/r2_agg_final -f python bench.py --steps 1 --warmup 0 --no-regimes --no-legs --no-cpu-baseline
timeout 300 ncu --set full --clock-control none -k regex:k_em_probs_tma -c 1 -o gpurun_out/r2_em_final -f python scripts/bench_em.py --cells 200000 --steps 1 --warmup 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_yg -c 1 -o gpurun_out/r2_yg_c4 -f python scripts/bench_g.py --cells 200000 --steps 1 --warmup 0 --flip 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_sweep_[yz]$' -c 2 -o gpurun_out/r2_sweeps_churn -f python bench.py --start random --steps 1 --warmup 0 --no-regimes --no-legs --no-cpu-baseline
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2_final.csv python bench.py --steps 2 --warmup 3 --no-regimes --no-legs --no-cpu-baseline
# two GPUs (gpurun --gpus 2): sharded EM + grid search legs                                                   # profiles/r2_bench_2gpu.json
# python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2
