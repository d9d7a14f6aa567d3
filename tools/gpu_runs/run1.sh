nvidia-smi > gpurun_out/gpu.txt 2>&1
tools/microbench/fp64_rates > gpurun_out/fp64_rates.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_golden.py -q 2>&1 | tail -40 > gpurun_out/parity1.log
timeout 200 python tools/loglik_profile.py 4096 32 > gpurun_out/ll4096.json 2>&1
timeout 200 python tools/loglik_profile.py 2048 20 > gpurun_out/ll2048.json 2>&1
timeout 200 python tools/loglik_profile.py 512 8 > gpurun_out/ll512.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_ll4096_r02.csv python tools/loglik_profile.py 4096 32 > /dev/null 2>&1
timeout 900 python tools/next_bench.py --workload headline --burnin 0 --calls 1 > gpurun_out/next_headline_r02_base.json 2>gpurun_out/next_headline_r02_base.err
echo done
