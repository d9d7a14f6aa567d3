for st in 2 3; do for B in 1 2 6; do SMK_LL_BG_STAGES=$st timeout 200 python tools/loglik_stages.py 4096 32 $B | cut -c1-130; done; done > gpurun_out/loglik_stages_r02d.txt 2>&1
echo done
