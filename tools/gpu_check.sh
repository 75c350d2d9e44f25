# Scratch script for one-off `gpurun -- bash tools/gpu_check.sh` calls.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "fp32" 2>&1 | tail -n 8 | cut -c 1-300
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
for pr in 64 32; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --precision $pr | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('precision $pr', round(d['value']), d['kernel_ms_per_step'], round(d['roofline']['frac'],3), d['config']['status_or'], d['clocks']['sm_mhz'])"; done
