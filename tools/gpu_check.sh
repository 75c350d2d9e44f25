cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for wl in c3 c2 c5 c5mpc; do python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/bench_${wl}_r1h.json 2> gpurun_out/bench_${wl}_r1h.err; python -c "
import json
d=json.load(open('gpurun_out/bench_${wl}_r1h.json'))
print('$wl', round(d['value']), round(d['e2e']['value']), d['kernel_ms_per_step'], round(d['roofline']['frac'],3), round(d['cpu_baseline']['value'],1), d['clocks']['sm_mhz'], d['config']['status_or'], d['gpu_launches'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench_r1h.log 2>&1; tail -n 1 gpurun_out/ncu_bench_r1h.log | cut -c 1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:egg_pgs_stream -c 1 -s 1 -o gpurun_out/prof_pgs_r1h python tools/profile_run.py c3 16384 20 2 > gpurun_out/ncu8.log 2>&1; tail -n 2 gpurun_out/ncu8.log | cut -c 1-200
