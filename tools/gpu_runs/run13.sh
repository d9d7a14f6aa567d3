timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_parity_at_size.py -q 2>&1 | tail -5 > gpurun_out/gputests_r02j.log
timeout 600 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_headline_r02g.json 2> gpurun_out/bench_headline_r02g.err
timeout 600 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_headline_s5_r02g.json 2> gpurun_out/bench_headline_s5_r02g.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:kxt_tc_kernel -c 4 --csv --log-file gpurun_out/kxt_tc_r02g.csv python bench.py --no-next --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
echo done
