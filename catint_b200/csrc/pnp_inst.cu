// pnp_inst.cu -- explicit instantiations of the block-size templated kernels, one block size per
// translation unit (compiled with -DCATINT_NB=<b>) so that the block sizes build in parallel.
// ST = false: default Poisson BCs (b = S+1); ST = true: Stern/Robin wall, phi carried (b = S+2).
#include "pnp_kernels.cuh"
#ifndef CATINT_NB
#error "compile with -DCATINT_NB=<block size>"
#endif
namespace catint {
template int launch_bdf<CATINT_NB, false>(SolveParams&, cudaStream_t);
template int launch_jac<CATINT_NB, false>(JacParams&, cudaStream_t);
#if CATINT_NB >= 3
template int launch_bdf<CATINT_NB, true>(SolveParams&, cudaStream_t);
#endif
}
