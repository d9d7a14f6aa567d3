timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gputests_r02u.log
for f in 0 1; do
SMK_FUSED_INVERSE=$f timeout 300 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_s5_fused$f.json 2> gpurun_out/bench_s5_fused$f.err
SMK_FUSED_INVERSE=$f timeout 300 python bench.py --no-next --no-cpu --steps 5 --samples 10 > gpurun_out/bench_s10_fused$f.json 2> gpurun_out/bench_s10_fused$f.err
SMK_FUSED_INVERSE=$f timeout 300 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_s40_fused$f.json 2> gpurun_out/bench_s40_fused$f.err
done
echo done
