timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_parity_at_size.py -q -x 2>&1 | tail -3 > gpurun_out/gputests_r02w.log
timeout 300 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_s40_sleep.json 2> gpurun_out/bench_s40_sleep.err
timeout 300 python bench.py --no-next --no-cpu --steps 3 > gpurun_out/bench_s40_sleep3.json 2> gpurun_out/bench_s40_sleep3.err
timeout 300 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_s5_sleep.json 2> gpurun_out/bench_s5_sleep.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:kxt_tc_kernel -c 4 --csv --log-file gpurun_out/kxt_tc_sleep.csv python bench.py --no-next --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
echo done
