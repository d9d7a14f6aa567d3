timeout 600 python -m pytest tests/test_gpu_golden.py -q -x -k "medium_n" 2>&1 | tail -30 > gpurun_out/gputests_r02n.log
timeout 600 python tools/parity_probe.py c2 4 2000 > gpurun_out/parity_c2_r02n.txt 2>&1
SMK_TC_MIN_N=0 timeout 600 python tools/parity_probe.py c2 4 2000 > gpurun_out/parity_c2_tc_r02n.txt 2>&1
echo done
