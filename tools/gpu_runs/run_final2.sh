timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_gputests.log
timeout 900 python bench.py > gpurun_out/final_bench_headline.json 2> gpurun_out/final_bench_headline.err
timeout 300 python bench.py --workload c4 --no-next > gpurun_out/final_bench_c4.json 2> gpurun_out/final_bench_c4.err
timeout 300 python bench.py --workload c2 --no-next > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
timeout 300 python bench.py --workload c3 --no-next > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err
timeout 300 python bench.py --no-next --no-cpu --samples 5 > gpurun_out/final_bench_s5.json 2> gpurun_out/final_bench_s5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_headline.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:predict_tc_kernel --csv --log-file gpurun_out/ncu_traffic_headline.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
echo done
