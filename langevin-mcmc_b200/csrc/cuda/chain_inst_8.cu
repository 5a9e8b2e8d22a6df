// chain_inst_8.cu -- k_chain_{init,run,stats}<8> (maxdepth <= 8)
#include "chain_kernels.cuh"
namespace lmc_cuda { LMC_INSTANTIATE_CHAIN(8) }
