for cfg in "tc tc" "simt tc" "simt simt"; do
  set -- $cfg
  SMK_FACTOR_IMPL=$1 SMK_PREDICT_IMPL=$2 timeout 300 python tools/parity_probe.py c2 10 4000 >> gpurun_out/parity_probe.jsonl 2>> gpurun_out/parity_probe.err
  SMK_FACTOR_IMPL=$1 SMK_PREDICT_IMPL=$2 timeout 300 python tools/parity_probe.py c4 4 4000 >> gpurun_out/parity_probe.jsonl 2>> gpurun_out/parity_probe.err
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests_r02a.log
timeout 200 python tools/loglik_profile.py 4096 32 > gpurun_out/ll4096_new.json 2>&1
timeout 200 python tools/loglik_profile.py 2048 20 > gpurun_out/ll2048_new.json 2>&1
timeout 200 python tools/loglik_profile.py 512 8 > gpurun_out/ll512_new.json 2>&1
SMK_LOGLIK_GRAPH=0 timeout 200 python tools/loglik_profile.py 4096 32 > gpurun_out/ll4096_new_nograph.json 2>&1
SMK_LOGLIK_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_ll4096_r02b.csv python tools/loglik_profile.py 4096 32 > /dev/null 2>&1
echo done
