# torchrun on N GPUs of one box: sharded-path parity (mgpu_check.py), then bench lines c2 / c3 / c5 [/ c4].   gpurun --gpus 8 -- "bash scripts/gpu_mgpu.sh 8 noc4"
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR scripts/mgpu_check.py > gpurun_out/mgpu_check_$N.log 2>&1; echo "mgpu check rc=$?"; grep -c "parity ok" gpurun_out/mgpu_check_$N.log; tail -2 gpurun_out/mgpu_check_$N.log
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; cat gpurun_out/bench_n$N.json | cut -c1-330; tail -2 gpurun_out/bench_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --workload c3 > gpurun_out/bench_c3_n$N.json 2> gpurun_out/bench_c3_n$N.err; echo "bench c3 N=$N rc=$?"; cat gpurun_out/bench_c3_n$N.json | cut -c1-330
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --workload c5 > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; echo "bench c5 N=$N rc=$?"; cat gpurun_out/bench_c5_n$N.json | cut -c1-330
if [ "${2:-}" != "noc4" ]; then timeout 900 $TR bench.py --gpus $N --steps 2 --warmup 3 --workload c4 --sites 4000000 --max-splits 256 > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; echo "bench c4 N=$N rc=$?"; cat gpurun_out/bench_c4_n$N.json | cut -c1-330; tail -3 gpurun_out/bench_c4_n$N.err; fi
