cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_r1g.json 2> gpurun_out/bench_n2_r1g.err
echo "rc=$?"; tail -c 600 gpurun_out/bench_n2_r1g.json; tail -n 15 gpurun_out/bench_n2_r1g.err
