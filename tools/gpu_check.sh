# Scratch script for one-off `gpurun -- bash tools/gpu_check.sh` calls.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1i_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench_r1i.log 2>&1; tail -n 1 gpurun_out/ncu_bench_r1i.log | cut -c 1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:egg_pgs_stream -c 1 -s 1 -o gpurun_out/prof_pgs_r1i python tools/profile_run.py c3 16384 20 2 > gpurun_out/ncu9.log 2>&1; tail -n 2 gpurun_out/ncu9.log | cut -c 1-200
