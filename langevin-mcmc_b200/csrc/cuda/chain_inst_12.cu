// chain_inst_12.cu -- k_chain_{init,run,stats}<12> (maxdepth <= 12)
#include "chain_kernels.cuh"
namespace lmc_cuda { LMC_INSTANTIATE_CHAIN(12) }
