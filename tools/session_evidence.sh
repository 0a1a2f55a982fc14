cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/pytest_r2.log
timeout 600 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench rc $?"; head -c 600 gpurun_out/bench_r2_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-tol > gpurun_out/ncu_l.log 2>&1; echo "ncu list rc $?"
ITERS=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_pass_kernel|tc_wside_kernel|tc_wfinish_kernel' -s 4 -c 6 -f -o gpurun_out/r2_iteration_full python tools/ncu_pass.py > gpurun_out/ncu_f.log 2>&1; echo "ncu full rc $?"; tail -3 gpurun_out/ncu_f.log
timeout 600 python tools/measure_configs.py cfg1 cfg2 cfg4 > gpurun_out/r2_other_configs.txt 2>&1; echo "configs rc $?"; cat gpurun_out/r2_other_configs.txt
