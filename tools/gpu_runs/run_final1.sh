# round-2 final measurements, one GPU
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/final_gputests.log
timeout 900 python bench.py > gpurun_out/final_bench_headline.json 2> gpurun_out/final_bench_headline.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
for wl in c2 c3 c4 c5; do timeout 600 python bench.py --workload $wl --no-next > gpurun_out/final_bench_$wl.json 2> gpurun_out/final_bench_$wl.err; done
timeout 300 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/final_bench_s5.json 2> gpurun_out/final_bench_s5.err
for B in 1 2 4 6; do timeout 200 python tools/loglik_stages.py 4096 32 $B; done > gpurun_out/final_loglik_stages.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_headline.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:predict_tc_kernel --csv --log-file gpurun_out/ncu_traffic_headline.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kxt_tc_kernel -s 1 -c 1 -o gpurun_out/kxt_tc_full_final -f python bench.py --no-next --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
echo done
