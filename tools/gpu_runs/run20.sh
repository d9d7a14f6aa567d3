timeout 120 tools/microbench/diag_bench > gpurun_out/diag_bench_r02j.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gputests_r02m.log
timeout 300 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_headline_r02j.json 2> gpurun_out/bench_headline_r02j.err
timeout 300 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_headline_s5_r02j.json 2> gpurun_out/bench_headline_s5_r02j.err
timeout 300 python tools/loglik_profile.py 4096 32 > gpurun_out/loglik_profile_r02j.txt 2>&1
echo done
