cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc $?"; tail -10 gpurun_out/pytest_r2b.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 240 python tools/exp_kl_cap.py > gpurun_out/r2_kl_variants.txt 2>&1; echo "kl variants rc $?"; cat gpurun_out/r2_kl_variants.txt
timeout 300 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc $?"; head -c 300 gpurun_out/bench_r2b.json
