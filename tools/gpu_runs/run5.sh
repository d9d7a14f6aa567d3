tools/microbench/diag_bench > gpurun_out/diag_bench3.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/gputests_r02d.log
for wl in "c2 10" "c4 4" "c3 2" "headline 2"; do
  set -- $wl
  timeout 600 python tools/parity_probe.py $1 $2 4000 >> gpurun_out/parity_probe3.jsonl 2>> gpurun_out/parity_probe3.err
done
timeout 200 python tools/loglik_profile.py 4096 32 > gpurun_out/ll4096_v4.json 2>&1
timeout 200 python tools/loglik_profile.py 2048 20 > gpurun_out/ll2048_v4.json 2>&1
timeout 600 python bench.py --no-next --steps 3 > gpurun_out/bench_headline_r02b.json 2> gpurun_out/bench_headline_r02b.err
timeout 600 python bench.py --no-next --no-cpu --steps 3 --samples 5 > gpurun_out/bench_headline_s5_r02b.json 2> gpurun_out/bench_headline_s5_r02b.err
timeout 600 python bench.py --no-next --workload c2 --steps 5 > gpurun_out/bench_c2_r02b.json 2> gpurun_out/bench_c2_r02b.err
timeout 600 python bench.py --no-next --workload c4 --steps 5 > gpurun_out/bench_c4_r02b.json 2> gpurun_out/bench_c4_r02b.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_headline_s5_r02b.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 --samples 5 > /dev/null 2>&1
SMK_LOGLIK_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_kernel -s 40 -c 2 -o gpurun_out/dgemm_r02 python tools/loglik_profile.py 4096 32 > /dev/null 2>&1
echo done
