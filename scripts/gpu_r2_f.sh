# Round-2 evidence run F (two B200s): sharded parity (scripts/mgpu_check.py) and the full bench line at N = 2.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/mgpu_check.py > gpurun_out/r2f_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; tail -4 gpurun_out/r2f_mgpu_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"; python scripts/bench_brief.py gpurun_out/r2f_bench_n2.json; tail -3 gpurun_out/r2f_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2f_bench_ref_n2.json 2> gpurun_out/r2f_bench_ref_n2.err; echo "ref n2 rc=$?"; cut -c1-400 gpurun_out/r2f_bench_ref_n2.json
