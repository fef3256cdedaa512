#!/bin/bash
# rebuild only the C-ABI object with K1 variants and time the RHS kernel (run on the GPU box)
set -e
cd "$(dirname "$0")/.."
OBJ=build/obj
for V in "-DCATINT_RHS_MINB=2" "-DCATINT_RHS_MINB=3" "-DCATINT_RHS_MINB=2 -DCATINT_RHS_NOREACT" $EXTRA_VARIANTS; do
  echo "=== variant: $V"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $V -c catint_b200/csrc/pnp_capi.cu -o /tmp/capi_var.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o catint_b200/libcatint_pnp.so $(ls $OBJ/pnp_inst_*.o) /tmp/capi_var.o
  python scripts/time_rhs.py 2>&1 | grep -v WARN | grep "GB/s"
done
# restore the default build
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o catint_b200/libcatint_pnp.so $(ls $OBJ/pnp_inst_*.o) $OBJ/pnp_capi.o
