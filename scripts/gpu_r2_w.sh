# Round-2 run W: the full bench line of the final tree at N = $1 GPUs.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2w_bench_n$N.json 2> gpurun_out/r2w_bench_n$N.err; echo "bench n$N rc=$?"; python scripts/bench_brief.py gpurun_out/r2w_bench_n$N.json 2>/dev/null | head -4; tail -2 gpurun_out/r2w_bench_n$N.err
