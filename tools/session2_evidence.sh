cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc $?"; tail -12 gpurun_out/pytest_r2b.log
timeout 240 python tools/exp_kl_cap.py > gpurun_out/r2_kl_cap.txt 2>&1; echo "kl cap rc $?"; cat gpurun_out/r2_kl_cap.txt
timeout 300 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc $?"; head -c 400 gpurun_out/bench_r2b.json
