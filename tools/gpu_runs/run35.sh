for ml in 64 48; do SMK_KXT_TC_MIN_LANES=$ml timeout 300 python bench.py --no-next --no-cpu --samples 20 > gpurun_out/bench_s20_ml$ml.json 2> gpurun_out/bench_s20_ml$ml.err; done
for ml in 64 24; do SMK_KXT_TC_MIN_LANES=$ml timeout 300 python bench.py --no-next --no-cpu --samples 10 > gpurun_out/bench_s10_ml$ml.json 2> gpurun_out/bench_s10_ml$ml.err; done
echo done
