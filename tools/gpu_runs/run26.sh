timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gputests_r02r.log
for B in 1 2; do timeout 200 python tools/loglik_stages.py 4096 32 $B; done > gpurun_out/loglik_stages_r02c.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_headline_full_r02k.json 2> gpurun_out/bench_headline_full_r02k.err
echo done
