tools/microbench/diag_bench > gpurun_out/diag_bench.txt 2>&1
timeout 300 python tools/tc_error_probe.py c2 4 2000 > gpurun_out/tc_error_probe_c2.txt 2>&1
timeout 300 python tools/tc_error_probe.py c4 1 2000 > gpurun_out/tc_error_probe_c4.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_chooser.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15 > gpurun_out/gputests_r02b.log
echo done
