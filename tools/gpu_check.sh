cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
bash tools/stress.sh 2>&1 | head -9
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_r1h.json 2> gpurun_out/bench_c3_r1h.err; python -c "
import json
d=json.load(open('gpurun_out/bench_c3_r1h.json'))
print(d['value'], d['e2e']['value'], d['kernel_ms_per_step'], d['roofline']['frac'], d['config']['status_or'])"
