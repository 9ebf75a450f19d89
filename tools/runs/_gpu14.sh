cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense.py -m gpu -x -q -s 2>&1 | tail -8 > gpurun_out/r02t_dense_tests.log
cat gpurun_out/r02t_dense_tests.log
timeout 600 python tools/dense_probe.py --batch 1024 2>&1 | tail -3 | tee gpurun_out/r02t_dense_probe.txt
