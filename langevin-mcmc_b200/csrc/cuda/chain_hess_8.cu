// chain_hess_8.cu -- the H2MC kernels for maxdepth <= 8: k_wave_grad<8, 2> (gradient + Hessian, forward-over-reverse)
// and k_h2mc_gaussian<8> (warp-cooperative ComputeGaussian), see chain_kernels.cuh LMC_INSTANTIATE_HESS.
#include "chain_kernels.cuh"
namespace lmc_cuda { LMC_INSTANTIATE_HESS(8) }
