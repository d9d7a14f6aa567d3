for wl in "c2 10" "c4 4" "c3 2" "headline 2" "c5 1"; do
  set -- $wl
  SMK_TC_GUARD=1e9 timeout 600 python tools/parity_probe.py $1 $2 4000 >> gpurun_out/parity_probe5.jsonl 2>> gpurun_out/parity_probe5.err
done
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_golden.py tests/test_gpu_kernels.py -q 2>&1 | tail -15 > gpurun_out/gputests_r02e.log
timeout 600 python bench.py --no-next --steps 3 > gpurun_out/bench_headline_r02c.json 2> gpurun_out/bench_headline_r02c.err
echo done
