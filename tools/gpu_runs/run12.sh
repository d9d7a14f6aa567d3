timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/gputests_r02i.log
for wl in "c3 2" "headline 2" "c5 1"; do
  set -- $wl
  timeout 600 python tools/parity_probe.py $1 $2 4000 >> gpurun_out/parity_probe7.jsonl 2>> gpurun_out/parity_probe7.err
done
# launch list of one headline step (shares of the step) and the DRAM traffic of the predict GEMM
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_headline_r02.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:predict_tc_kernel --csv --log-file gpurun_out/ncu_traffic_headline.csv python bench.py --no-next --no-cpu --steps 1 --warmup 3 > /dev/null 2>&1
# full captures: predict GEMM (mode 0, one full chunk), Cholesky update (mode 2), generator, DMMA GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:predict_tc_kernel -s 60 -c 3 -o gpurun_out/predict_tc_r02 python bench.py --no-next --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kxt_tc_kernel -c 1 -o gpurun_out/kxt_tc_r02 python bench.py --no-next --no-cpu --steps 1 --warmup 1 > /dev/null 2>&1
SMK_LOGLIK_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgemm_nt_kernel -s 4 -c 2 -o gpurun_out/dgemm_r02b python tools/loglik_profile.py 4096 32 > /dev/null 2>&1
echo done
