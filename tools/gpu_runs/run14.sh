timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/gputests_r02k.log
timeout 300 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_headline_s5_r02h.json 2> gpurun_out/bench_headline_s5_r02h.err
timeout 900 python bench.py > gpurun_out/bench_headline_full_r02h.json 2> gpurun_out/bench_headline_full_r02h.err
echo done
