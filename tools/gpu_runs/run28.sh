timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mlii.py -q -x 2>&1 | tail -3 > gpurun_out/gputests_r02s.log
for pdl in 0 1; do for B in 1 2 6; do SMK_LL_PDL=$pdl timeout 200 python tools/loglik_stages.py 4096 32 $B | cut -c1-130; done; done > gpurun_out/loglik_stages_r02e.txt 2>&1
SMK_LL_PDL=1 timeout 200 python tools/loglik_stages.py 2048 20 1 | cut -c1-130 >> gpurun_out/loglik_stages_r02e.txt 2>&1
SMK_LL_PDL=1 SMK_LOGLIK_GRAPH=0 timeout 200 python tools/loglik_stages.py 4096 32 1 | cut -c1-130 >> gpurun_out/loglik_stages_r02e.txt 2>&1
echo done
