cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/compare_variants.py c3 64 50 3 stream fast | tail -n 2
timeout 300 python tools/compare_variants.py c2 67 500 2 stream fast | tail -n 2
timeout 300 python tools/compare_variants.py c5 61 500 2 stream:1 fast | tail -n 2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_run.py c3 16384 20 3 2>&1 | tail -n 1; }
run EGG_PGS_PF=0 EGG_PGS_PFMODE=2
run EGG_PGS_PF=0 EGG_PGS_PFMODE=0
run EGG_PGS_PF=3 EGG_PGS_PFMODE=2
run EGG_PGS_PF=0 EGG_PGS_PFMODE=2 EGG_PGS_CTAS_PER_SM=8
run EGG_PGS_PF=0 EGG_PGS_PFMODE=2 EGG_PGS_ISO=1
