timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gputests_r02o.log
timeout 300 python tools/loglik_profile.py 4096 32 > gpurun_out/loglik_profile_r02k.txt 2>&1
timeout 300 python tools/loglik_profile.py 2048 20 >> gpurun_out/loglik_profile_r02k.txt 2>&1
timeout 300 python tools/loglik_profile.py 512 8 >> gpurun_out/loglik_profile_r02k.txt 2>&1
echo done
