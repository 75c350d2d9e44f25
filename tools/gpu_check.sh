# Scratch script for one-off `gpurun -- bash tools/gpu_check.sh` calls: tests, smoke and a short bench.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c 1-200
python bench.py --steps 2 --warmup 3 --no-cpu-baseline | cut -c 1-400
