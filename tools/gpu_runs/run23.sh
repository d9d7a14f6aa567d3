timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "potrf or loglik or chol" 2>&1 | tail -3 > gpurun_out/gputests_r02p.log
timeout 300 python tools/loglik_profile.py 4096 32 > gpurun_out/loglik_profile_r02l.txt 2>&1
timeout 300 python tools/loglik_profile.py 2048 20 >> gpurun_out/loglik_profile_r02l.txt 2>&1
echo done
