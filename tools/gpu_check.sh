# Scratch script for one-off `gpurun -- bash tools/gpu_check.sh` calls: tests, smoke, stress and a short bench.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
bash tools/stress.sh 2>&1 | cut -c 1-150
