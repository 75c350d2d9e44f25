cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 12 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_run.py c3 16384 20 3 2>&1 | tail -n 1; }
run A=1
