for B in 1 2 6; do timeout 200 python tools/loglik_stages.py 4096 32 $B; done > gpurun_out/loglik_stages_r02.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k regex:'diag_kernel|panel_kernel|dgemm_nt' -s 300 -c 330 --csv --log-file gpurun_out/ll_launches_r02.csv python tools/loglik_stages.py 4096 32 1 > gpurun_out/ll_ncu.log 2>&1
echo done
