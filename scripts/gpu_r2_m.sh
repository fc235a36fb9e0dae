# Round-2 run M (two B200s): sharded parity and the full bench line at N = 2 with the cost-balanced split partition.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 scripts/mgpu_check.py > gpurun_out/r2m_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; tail -3 gpurun_out/r2m_mgpu_check.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo "bench n2 rc=$?"; python scripts/bench_brief.py gpurun_out/r2m_bench_n2.json 2>/dev/null; tail -2 gpurun_out/r2m_bench_n2.err
