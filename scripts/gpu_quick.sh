cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 -x -k "score or scorer or config2 or flat" > gpurun_out/t_quick.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/t_quick.log
timeout 600 python scripts/phase_profile.py > gpurun_out/phase.log 2>&1; echo "phase rc=$?"; cat gpurun_out/phase.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
python scripts/ncu_kernels.py > gpurun_out/ncu_kernels_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"count_kernel|u8_kernel|pair_kernel|pack_kernel|symv_block|subflatten_score|score_small|krylov_rr|gram_u8_small|hi_cross" -c 40 -o gpurun_out/prof_kernels python scripts/ncu_kernels.py > gpurun_out/ncu_kernels.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_kernels.log; ls -la gpurun_out/prof_kernels.ncu-rep
