#!/bin/bash
# round-2 profile artefacts (run under gpurun; outputs in gpurun_out/)
set -x
# (1) every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
# (2) DRAM traffic of the lock-step launches of one sweep
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dense_ --csv \
    --log-file gpurun_out/r2_lockstep_dram.csv python scripts/prof_sweep.py c2 1 > /dev/null 2>&1
# (3) ncu --set full of the lock-step tile passes and of the 2-warp register-tile kernel (largest regtile bin)
scripts/ncu_kernel.sh 'dense_walk_kernel' r2_lockstep_walk 3 python scripts/prof_sweep.py c2 1
scripts/ncu_kernel.sh 'regtile_kernel<.int.4, .int.3, .int.2, .int.2>' r2_regtile2w48 2 python scripts/prof_sweep.py c2 1
