timeout 300 python tools/tc_error_probe.py c2 4 2000 > gpurun_out/tc_error_probe2_c2.txt 2>&1
timeout 300 python tools/tc_error_probe.py c4 1 2000 > gpurun_out/tc_error_probe2_c4.txt 2>&1
timeout 300 python tools/tc_error_probe.py headline 1 1500 > gpurun_out/tc_error_probe2_hl.txt 2>&1
