timeout 900 ncu --set full --clock-control none --import-source on -k regex:kxt_tc_kernel -s 1 -c 1 -o gpurun_out/kxt_tc_full_r02h -f python bench.py --no-next --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_kxt.log 2>&1
echo done
