# One-stop GPU check: `gpurun -- bash tools/gpu_check.sh` (tests, smoke, reproducibility, stress, short bench).
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | cut -c 1-160
timeout 600 python tools/determinism_check.py c3 16384 20 3 2>&1 | tail -n 1
bash tools/stress.sh 2>&1 | head -9 | cut -c 1-150
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c 1-300
