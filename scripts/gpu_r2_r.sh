# Round-2 run R (two B200s): bench at N = 2, refit with the fixed-cost scale (warm-up 3 = one re-cut, warm-up 5 = three).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; echo "bench n2 rc=$?"; python scripts/bench_brief.py gpurun_out/r2r_bench_n2.json 2>/dev/null; tail -2 gpurun_out/r2r_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29682 bench.py --gpus 2 --steps 5 --warmup 5 --extras none > gpurun_out/r2r_bench_n2_w5.json 2> gpurun_out/r2r_bench_n2_w5.err; echo "bench n2 (warmup 5) rc=$?"; python scripts/bench_brief.py gpurun_out/r2r_bench_n2_w5.json 2>/dev/null | grep -v roofline
