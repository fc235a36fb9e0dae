# Round-2 run P (two B200s): bench at N = 2 with the one-shot partition refit (against SPB_PARTITION_REFIT=0).
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 5 --warmup 3 --extras none > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err; echo "bench n2 rc=$?"; python scripts/bench_brief.py gpurun_out/r2p_bench_n2.json 2>/dev/null; tail -2 gpurun_out/r2p_bench_n2.err
SPB_PARTITION_REFIT=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus 2 --steps 5 --warmup 3 --extras none > gpurun_out/r2p_bench_n2_norefit.json 2> gpurun_out/r2p_bench_n2_norefit.err; echo "bench n2 (no refit) rc=$?"; python scripts/bench_brief.py gpurun_out/r2p_bench_n2_norefit.json 2>/dev/null | head -3
