cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_run.py c3 16384 20 3 2>&1 | tail -n 1; }
run EGG_PGS_PF_SPAN=0
run EGG_PGS_PF_SPAN=11
run EGG_PGS_PF_SPAN=12
run EGG_PGS_PF_SPAN=13
run EGG_PGS_PF_SPAN=14
run EGG_PGS_PF_SPAN=13 EGG_PGS_CTAS_PER_SM=8
timeout 300 python tools/compare_variants.py c3 64 50 3 stream fast | tail -n 2
