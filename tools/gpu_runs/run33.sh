timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/gputests_r02x.log
for i in 1 2 3 4; do timeout 300 python bench.py --no-next --no-cpu > gpurun_out/bench_var$i.json 2> gpurun_out/bench_var$i.err; done
timeout 300 python bench.py --no-next --no-cpu --samples 5 > gpurun_out/bench_var_s5.json 2> gpurun_out/bench_var_s5.err
timeout 300 python bench.py --no-next --no-cpu --workload c4 > gpurun_out/bench_var_c4.json 2> gpurun_out/bench_var_c4.err
echo done
