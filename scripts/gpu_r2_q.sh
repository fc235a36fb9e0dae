# Round-2 run Q (eight B200s): the full bench line at N = 8 with the one-shot partition refit.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err; echo "bench n8 rc=$?"; python scripts/bench_brief.py gpurun_out/r2q_bench_n8.json 2>/dev/null; tail -2 gpurun_out/r2q_bench_n8.err
