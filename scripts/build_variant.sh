#!/bin/bash
# Variant build of the K3 instance for ONE block size (default 9 = C2) with extra compiler flags; the other objects
# come from build/obj (run __graft_entry__.build() first).  For A/B measurements:
#   scripts/build_variant.sh "-DCATINT_L2_HINTS=0" build/variants/nohint.so [NB]
#   CATINT_PNP_LIB=build/variants/nohint.so python scripts/profile_case.py ...
set -e
cd "$(dirname "$0")/.."
FLAGS="$1"; OUT="$2"; NB="${3:-9}"
mkdir -p "$(dirname "$OUT")" build/variants
OBJ=build/variants/inst_${NB}_$(echo "$FLAGS" | md5sum | cut -c1-8).o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $FLAGS -DCATINT_NB=$NB -c catint_b200/csrc/pnp_inst.cu -o $OBJ
OTHERS=$(ls build/obj/pnp_inst_*.o | grep -v "pnp_inst_${NB}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $OTHERS $OBJ build/obj/pnp_capi.o
echo "built $OUT"
