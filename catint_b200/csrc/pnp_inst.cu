// pnp_inst.cu -- one explicit instantiation of the block-size templated kernels per
// translation unit (compiled with -DCATINT_NB=<b>), so that the block sizes build in parallel.
#include "pnp_kernels.cuh"
#ifndef CATINT_NB
#error "compile with -DCATINT_NB=<block size>"
#endif
namespace catint {
template int launch_bdf<CATINT_NB>(SolveParams&, cudaStream_t);
template int launch_jac<CATINT_NB>(JacParams&, cudaStream_t);
}
