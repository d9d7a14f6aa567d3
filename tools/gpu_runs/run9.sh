timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/gputests_r02f.log
for wl in "c2 10" "c4 4" "c3 2" "headline 2" "c5 1"; do
  set -- $wl
  timeout 600 python tools/parity_probe.py $1 $2 4000 >> gpurun_out/parity_probe6.jsonl 2>> gpurun_out/parity_probe6.err
done
timeout 900 python bench.py --steps 3 > gpurun_out/bench_headline_r02d.json 2> gpurun_out/bench_headline_r02d.err
timeout 600 python bench.py --no-next --no-cpu --steps 3 --samples 5 > gpurun_out/bench_headline_s5_r02d.json 2> gpurun_out/bench_headline_s5_r02d.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_headline_s5_r02d.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 --samples 5 > /dev/null 2>&1
echo done
