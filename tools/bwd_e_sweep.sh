#!/bin/bash
# First GPU call of the next round: the unmeasured switches of infonce_bwd_e_kernel on the 64 distillation pairs at
# b = N = 16384 (bwd_e_check.py ... t = time the stored-exponential backward only).
#   gpurun --timeout 400 -- 'bash tools/bwd_e_sweep.sh > gpurun_out/bwd_e_sweep.log 2>&1; cat gpurun_out/bwd_e_sweep.log'
run() { echo "== $*"; env "$@" timeout 100 python tools/bwd_e_check.py 16384 16384 16 4 14.2857 t 2>&1 | tail -1; }
run COSMOS_B200_EPREFETCH=tensor
run COSMOS_B200_EPREFETCH=bulk
run COSMOS_B200_EPREFETCH=bulk COSMOS_B200_EAHEAD=3
run COSMOS_B200_EPREFETCH=bulk COSMOS_B200_EAHEAD=12
run COSMOS_B200_EAHEAD=0
echo "== parity of the bulk prefetch (ragged, scale 100)"
COSMOS_B200_EPREFETCH=bulk timeout 100 python tools/bwd_e_check.py 1000 1016 2 2 100 2>&1 | tail -8
