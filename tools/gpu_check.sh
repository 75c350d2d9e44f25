cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dense or host_mirror or stabilize" > gpurun_out/pytest_gpu_dense.log 2>&1; tail -n 3 gpurun_out/pytest_gpu_dense.log
timeout 600 python tools/dense_run.py 1024 3 2>&1 | tail -n 4
