// chain_inst_4.cu -- k_chain_{init,run,stats}<4> (maxdepth <= 4)
#include "chain_kernels.cuh"
namespace lmc_cuda { LMC_INSTANTIATE_CHAIN(4) }
