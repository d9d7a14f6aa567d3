timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mlii.py -q -x 2>&1 | tail -3 > gpurun_out/gputests_r02t.log
for B in 1 2 4 6; do timeout 200 python tools/loglik_stages.py 4096 32 $B | cut -c1-130; done > gpurun_out/loglik_stages_r02f.txt 2>&1
timeout 200 python tools/loglik_stages.py 2048 20 1 | cut -c1-130 >> gpurun_out/loglik_stages_r02f.txt 2>&1
timeout 200 python tools/loglik_stages.py 1000 8 1 | cut -c1-130 >> gpurun_out/loglik_stages_r02f.txt 2>&1
timeout 200 python tools/loglik_stages.py 8192 32 1 | cut -c1-130 >> gpurun_out/loglik_stages_r02f.txt 2>&1
echo done
