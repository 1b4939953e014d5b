// The host shim (libaffhost.so: pure C++, no CUDA) reaches the CUDA library (libaffpredict.so,
// include/affpredict.h) through this table, resolved on first use.  Candidate generation, scene
// editing and message formatting therefore work - and the CPU reference arm of bench.py runs -
// without the CUDA library being mapped; anything that predicts fails loudly if it is missing
// (there is no CPU fallback).
#ifndef AFF_HOST_GPU_API_H
#define AFF_HOST_GPU_API_H

#include "affpredict.h"

namespace affgpu
{

struct GpuApi
{
  decltype(&::aff_last_error) last_error;
  decltype(&::aff_scene_create) scene_create;
  decltype(&::aff_scene_destroy) scene_destroy;
  decltype(&::aff_scene_num_ik_joints) scene_num_ik_joints;
  decltype(&::aff_batch_create) batch_create;
  decltype(&::aff_batch_destroy) batch_destroy;
  decltype(&::aff_batch_launch) batch_launch;
  decltype(&::aff_batch_results) batch_results;
  decltype(&::aff_batch_num_steps) batch_num_steps;
  decltype(&::aff_batch_fetch_q) batch_fetch_q;
  decltype(&::aff_predict_batch) predict_batch;
};

// dlopen("libaffpredict.so" next to libaffhost.so) + dlsym of every entry; throws std::runtime_error
// naming the missing library or symbol.  loaded(): true once the table is resolved (no side effect).
const GpuApi& gpu();
bool gpuLoaded();

}   // namespace affgpu

#endif
