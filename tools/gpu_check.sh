# Scratch script for one-off `gpurun -- bash tools/gpu_check.sh` calls.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/n2.json; wc -l gpurun_out/n2.json; head -c 120 gpurun_out/n2.json; echo
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | wc -l
