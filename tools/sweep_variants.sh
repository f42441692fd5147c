for v in "$@"; do DDM_B200_LIB=/root/repo/build/variants/lib_$v.so python tools/time_sim.py 33554432 2>&1 | tail -1; done > gpurun_out/sweep.log 2>&1
