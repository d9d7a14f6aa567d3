tools/microbench/diag_bench > gpurun_out/diag_bench2.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/gputests_r02c.log
for wl in "c2 10" "c4 4" "c3 2" "headline 2" "c5 1"; do
  set -- $wl
  SMK_TC_GUARD=0 timeout 600 python tools/parity_probe.py $1 $2 4000 >> gpurun_out/parity_probe2.jsonl 2>> gpurun_out/parity_probe2.err
  timeout 600 python tools/parity_probe.py $1 $2 4000 >> gpurun_out/parity_probe2.jsonl 2>> gpurun_out/parity_probe2.err
done
timeout 200 python tools/loglik_profile.py 4096 32 > gpurun_out/ll4096_v3.json 2>&1
timeout 200 python tools/loglik_profile.py 2048 20 > gpurun_out/ll2048_v3.json 2>&1
timeout 200 python tools/loglik_profile.py 512 8 > gpurun_out/ll512_v3.json 2>&1
timeout 600 python bench.py --no-next --steps 3 > gpurun_out/bench_headline_r02a.json 2> gpurun_out/bench_headline_r02a.err
timeout 600 python bench.py --no-next --no-cpu --steps 3 --samples 5 > gpurun_out/bench_headline_s5_r02a.json 2> gpurun_out/bench_headline_s5_r02a.err
echo done
