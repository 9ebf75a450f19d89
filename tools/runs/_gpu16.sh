cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "matrix_game or qsaddle" 2>&1 | grep -v "tenscalc_b200: module" | tail -25 > gpurun_out/r02p_eq_tests.log
cat gpurun_out/r02p_eq_tests.log
