timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity_at_size.py -q 2>&1 | tail -5 > gpurun_out/gputests_r02h.log
timeout 600 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_headline_r02f.json 2> gpurun_out/bench_headline_r02f.err
SMK_KXT_IMPL=tc timeout 600 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_headline_r02f_kxttc.json 2> gpurun_out/bench_headline_r02f_kxttc.err
echo done
