for wl in "c2 10" "c4 4" "c3 2" "headline 2" "c5 1"; do
  set -- $wl
  SMK_TC_GUARD=1e9 timeout 600 python tools/parity_probe.py $1 $2 4000 >> gpurun_out/parity_probe4.jsonl 2>> gpurun_out/parity_probe4.err
done
tools/microbench/diag_bench > gpurun_out/diag_bench4.txt 2>&1
timeout 200 python tools/loglik_profile.py 4096 32 > gpurun_out/ll4096_v5.json 2>&1
timeout 200 python tools/ei_sweep_bench.py > gpurun_out/ei_sweep_r02.jsonl 2>&1
echo done
