// Entry points of include/affpredict.h that need no device: compiled into BOTH libaffpredict.so
// (the CUDA library) and libaffhost.so (the pure C++ host shim), so that the shim - and the CPU
// reference arm of bench.py - never needs the CUDA library for them.
#include <algorithm>

#include "affpredict.h"

extern "C" {

int aff_abi_version(void) { return AFF_ABI_VERSION; }

void aff_default_opts(aff_opts* o)
{
  if (!o) return;
  o->dt = 0.01; o->alpha = 0.05; o->lambda = 1.0e-6; o->q_filt_decay = 0.1; o->n_after_steps = 500; o->log_q = 0;
  o->early_exit = 0; o->reserved = 0;
}

void aff_rank_results(const aff_result* res, size_t n, int32_t* order)
{
  for (size_t i = 0; i < n; ++i) order[i] = (int32_t)i;
  std::stable_sort(order, order + n, [&](int32_t a, int32_t b) {
    const aff_result& x = res[a]; const aff_result& y = res[b];
    if ((x.success != 0) != (y.success != 0)) return x.success != 0;
    const double qx = x.jl_cost + x.coll_cost, qy = y.jl_cost + y.coll_cost;
    if (qx != qy) return qx < qy;
    return x.idx < y.idx;
  });
}

}   // extern "C"
