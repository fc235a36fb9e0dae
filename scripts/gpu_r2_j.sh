# Round-2 run J (eight B200s): sharded parity, the full bench line at N = 8, one vs two scoring streams.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 scripts/mgpu_check.py > gpurun_out/r2j_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; tail -3 gpurun_out/r2j_mgpu_check.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err; echo "bench n8 rc=$?"; python scripts/bench_brief.py gpurun_out/r2j_bench_n8.json 2>/dev/null; tail -2 gpurun_out/r2j_bench_n8.err
SPB_SCORE_STREAMS=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 8 --steps 5 --warmup 3 --extras none > gpurun_out/r2j_bench_n8_2s.json 2> gpurun_out/r2j_bench_n8_2s.err; echo "bench n8 (2 streams) rc=$?"; python scripts/bench_brief.py gpurun_out/r2j_bench_n8_2s.json 2>/dev/null | head -3
