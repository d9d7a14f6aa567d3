tools/microbench/diag_bench > gpurun_out/diag_bench5.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/gputests_r02g.log
timeout 600 python bench.py --no-next --steps 5 > gpurun_out/bench_headline_r02e.json 2> gpurun_out/bench_headline_r02e.err
SMK_FACTOR_OVERLAP=0 timeout 600 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_headline_r02e_noovl.json 2> gpurun_out/bench_headline_r02e_noovl.err
timeout 600 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_headline_s5_r02e.json 2> gpurun_out/bench_headline_s5_r02e.err
SMK_FACTOR_OVERLAP=0 timeout 600 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_headline_s5_r02e_noovl.json 2> gpurun_out/bench_headline_s5_r02e_noovl.err
timeout 200 python tools/loglik_profile.py 4096 32 > gpurun_out/ll4096_v6.json 2>&1
timeout 200 python tools/ei_sweep_bench.py > gpurun_out/ei_sweep_r02b.jsonl 2>&1
echo done
