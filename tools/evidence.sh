# Round-end evidence on one B200: `gpurun -- bash tools/evidence.sh <tag>` (tests, benches of every
# workload, ncu launch list of the default bench, one full-size ncu capture of the solve kernel).
cd ${GRAFT_REPO_ROOT:-.}
tag=${1:-rX}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -n 4 > gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 > gpurun_out/${tag}_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
for wl in c2 c4 c5 c5mpc; do python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/${tag}_bench_$wl.json 2> gpurun_out/${tag}_bench_$wl.err; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_c3_reference.json 2> gpurun_out/${tag}_bench_c3_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fixed-k --no-evolving > gpurun_out/${tag}_launches_bench.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:egg_pgs_stream_kernel -s 1 -c 1 -o gpurun_out/${tag}_pgs_stream_c3_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fixed-k --no-evolving > gpurun_out/${tag}_ncu_full.out 2>&1
ls -la gpurun_out | grep ${tag}_ | wc -l
