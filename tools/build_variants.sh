#!/bin/bash
# Kernel-variant sweeps of the simulator: build/variants/lib_<name>.so = csrc/ddm_sim.cu (or a file given as SRC=...)
# compiled with the extra -D flags of the variant, linked with the other translation units built once into build/obj.
#   tools/build_variants.sh name1 "-DDDM_SIM_THREADS=224 -DDDM_SIM_MIN_BLOCKS=4" name2 "..." ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/sbi_for_diffusion_models_b200/csrc
OBJ=$ROOT/build/obj; VAR=$ROOT/build/variants
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden"
mkdir -p $OBJ $VAR
for f in ddm_common ddm_pulses mnle_grad mnle_sampler mnle_simt mnle_tc mnle_train; do
  [ $OBJ/$f.o -nt $CS/$f.cu ] || nvcc $FL -c $CS/$f.cu -o $OBJ/$f.o &
done
[ $OBJ/ddm_pack_host.o -nt $CS/ddm_pack_host.cpp ] || nvcc $FL -c $CS/ddm_pack_host.cpp -o $OBJ/ddm_pack_host.o &
wait
SRC=${SRC:-$CS/ddm_sim.cu}
while [ $# -gt 0 ]; do
  name=$1; defs=$2; shift 2
  ( nvcc $FL $defs -I$CS -c $SRC -o $OBJ/sim_$name.o -Xptxas -v 2> $VAR/$name.ptxas.txt
    nvcc -shared -o $VAR/lib_$name.so $OBJ/sim_$name.o $OBJ/ddm_common.o $OBJ/ddm_pulses.o $OBJ/mnle_grad.o $OBJ/mnle_sampler.o \
         $OBJ/mnle_simt.o $OBJ/mnle_tc.o $OBJ/mnle_train.o $OBJ/ddm_pack_host.o
    echo "$name: $(grep -A2 'sim_kernelILi3ELb0ELb1ELi[0-9]ELb0ELb0' $VAR/$name.ptxas.txt | grep -o 'Used [0-9]* registers' | head -1) $(grep -A2 'sim_kernelILi3ELb0ELb1ELi[0-9]ELb0ELb0' $VAR/$name.ptxas.txt | grep -o '[0-9]* bytes spill stores' | head -1)" ) &
done
wait
