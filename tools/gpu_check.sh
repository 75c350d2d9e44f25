set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/compare_variants.py c3 64 50 3 > gpurun_out/cmp_c3.log 2>&1; echo "rc=$?" >> gpurun_out/cmp_c3.log
timeout 300 python tools/compare_variants.py c2 64 500 3 > gpurun_out/cmp_c2.log 2>&1; echo "rc=$?" >> gpurun_out/cmp_c2.log
timeout 300 python tools/compare_variants.py c5 64 500 3 > gpurun_out/cmp_c5.log 2>&1; echo "rc=$?" >> gpurun_out/cmp_c5.log
tail -5 gpurun_out/cmp_c3.log gpurun_out/cmp_c2.log gpurun_out/cmp_c5.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
for v in fast stream; do EGG_PGS_VARIANT=$v timeout 300 python tools/profile_run.py c3 8192 20 3 > gpurun_out/prof_$v.log 2>&1; tail -3 gpurun_out/prof_$v.log; done
