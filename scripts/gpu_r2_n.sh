# Round-2 run N (eight B200s): sharded parity and the full bench line at N = 8 with the cost-balanced split partition.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 scripts/mgpu_check.py > gpurun_out/r2n_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; tail -3 gpurun_out/r2n_mgpu_check.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n_bench_n8.json 2> gpurun_out/r2n_bench_n8.err; echo "bench n8 rc=$?"; python scripts/bench_brief.py gpurun_out/r2n_bench_n8.json 2>/dev/null; tail -2 gpurun_out/r2n_bench_n8.err
