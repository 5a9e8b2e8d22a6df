// pathgrad.h -- STUB (replaced by the hand-derived adjoint): gradient of log ssScore w.r.t. PSS.
#pragma once
#include "path.h"
namespace lmc {
template <int MAXD> LMC_HD bool grad_supported(const Scene &, const Path<MAXD> &) { return false; }
template <int MAXD> LMC_HD void path_gradient(const Scene &, const Path<MAXD> &, float *) {}
}
