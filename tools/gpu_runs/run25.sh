timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "potrf or loglik or chol" 2>&1 | tail -3 > gpurun_out/gputests_r02q.log
for B in 1 2 4 6; do timeout 200 python tools/loglik_stages.py 4096 32 $B; done > gpurun_out/loglik_stages_r02b.txt 2>&1
timeout 200 python tools/loglik_stages.py 2048 20 1 >> gpurun_out/loglik_stages_r02b.txt 2>&1
timeout 200 python tools/loglik_stages.py 8192 32 1 >> gpurun_out/loglik_stages_r02b.txt 2>&1
echo done
