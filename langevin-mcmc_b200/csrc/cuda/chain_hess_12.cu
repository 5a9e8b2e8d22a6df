// chain_hess_12.cu -- H2MC is limited to maxdepth <= 8 (dense Gaussians are stored for dim <= 16; lmc_create rejects
// h2mc with a larger maxdepth, as the reference has no derivative functions beyond path length 8, src/main.cpp:46), so the
// maxdepth <= 12 variant of the chain loop never launches the Hessian kernels: no instantiation, just the symbol.
#include "chain_kernels.cuh"
namespace lmc_cuda {
template <> cudaError_t launch_wave_hess<12>(cudaStream_t, const Scene &, ChainRec<12> *, int, const int *, const int *, int, H2mcSide *,
                                             H2mcSide *, int, int) {
    return cudaErrorNotSupported;
}
}  // namespace lmc_cuda
