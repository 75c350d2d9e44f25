# Scratch script for one-off `gpurun -- bash tools/gpu_check.sh` calls.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
for wl in c3 c2 c5 c5mpc; do python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/bench_${wl}_r1i.json 2> gpurun_out/bench_${wl}_r1i.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_${wl}_r1i.json') if l.startswith('{')][-1])
print('$wl', round(d['value']), round(d['e2e']['value']), d['kernel_ms_per_step'], round(d['roofline']['frac'],3), round(d['cpu_baseline']['value'],1), d['clocks']['sm_mhz'], d['config']['status_or'])"; done
