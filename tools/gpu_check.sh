# Scratch script for one-off `gpurun -- bash tools/gpu_check.sh` calls.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
det() { echo "== det $*: $(env $ENVV timeout 600 python tools/determinism_check.py "$@" 2>&1 | grep "^run\|DETERM" | tail -n 2 | tr '\n' ' ' | cut -c 1-170)"; }
ENVV="EGG_PGS_LPW=1" det c2 65536 20 3
ENVV="EGG_PGS_LPW=1" det c5 131072 20 3
ENVV="EGG_PGS_LPW=2" det c2 65536 20 3
ENVV="A=1" det c2 65536 20 3
ENVV="A=1" det c3 16384 20 3
ENVV="A=1" det c3 65536 5 3
ENVV="A=1" det c5 131072 50 3
ENVV="EGG_PGS_LPW=16" det c3 8192 10 3
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
