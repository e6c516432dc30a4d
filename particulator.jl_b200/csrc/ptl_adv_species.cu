// ptl_adv_species.cu — the advance kernels of ONE species (-DPTL_TU_SPECIES=0..3): compiled four times into separate
// objects so that the heavy template instantiations (species x table kind x first pass x callbacks x kernel variant)
// build in parallel.
#include "ptl_launch.cuh"

#ifndef PTL_TU_SPECIES
#error "compile with -DPTL_TU_SPECIES=<species id>"
#endif

namespace ptl_host {
template int32_t launch_advance_s<PTL_TU_SPECIES>(ptl_context*, const ptl::AdvanceParams&, long long, long long, bool, bool, size_t, bool);
}
