timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/gputests_r02v.log
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-next --no-cpu --steps 5 --samples 5 > gpurun_out/bench_s5_$tag.json 2> gpurun_out/bench_s5_$tag.err; env "$@" timeout 300 python bench.py --no-next --no-cpu --steps 5 > gpurun_out/bench_s40_$tag.json 2> gpurun_out/bench_s40_$tag.err; }
run legacy SMK_MEAN_FROM_GEMM=0
run zmean SMK_PREGEN=0
run pregen SMK_PREGEN=1
for wl in c2 c4 headline c5; do timeout 600 python tools/parity_probe.py $wl 4 2000 2>&1 | tail -1; done > gpurun_out/parity_r02v.txt
echo done
