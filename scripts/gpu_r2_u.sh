# Round-2 run U (four B200s): the full bench line at N = 4.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29691 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2u_bench_n4.json 2> gpurun_out/r2u_bench_n4.err; echo "bench n4 rc=$?"; python scripts/bench_brief.py gpurun_out/r2u_bench_n4.json 2>/dev/null; tail -2 gpurun_out/r2u_bench_n4.err
