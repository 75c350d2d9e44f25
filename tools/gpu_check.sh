cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
bash tools/stress.sh 2>&1 | head -9
for m in 0 16; do
echo "== pfmode $m"
EGG_PGS_PFMODE=$m timeout 300 python tools/profile_run.py c5 131072 500 2 2>&1 | tail -n 1
EGG_PGS_PFMODE=$m timeout 300 python tools/profile_run.py c2 65536 500 2 2>&1 | tail -n 1
EGG_PGS_PFMODE=$m timeout 300 python tools/profile_run.py c3 16384 20 3 2>&1 | tail -n 1
done
