cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "broadphase or box_bounds" > gpurun_out/pytest_gpu_new.log 2>&1; tail -n 8 gpurun_out/pytest_gpu_new.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_run.py c3 16384 20 3 2>&1 | tail -n 1
