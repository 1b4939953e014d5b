#include "gpu_api.h"

#include <dlfcn.h>

#include <mutex>
#include <stdexcept>
#include <string>

namespace affgpu
{

static GpuApi g_api;
static bool g_loaded = false;
static std::mutex g_mtx;

bool gpuLoaded() { return g_loaded; }

const GpuApi& gpu()
{
  std::lock_guard<std::mutex> lock(g_mtx);
  if (g_loaded) return g_api;
  // the CUDA library sits next to this one
  std::string dir;
  Dl_info info;
  if (dladdr((const void*)&gpuLoaded, &info) && info.dli_fname) {
    dir = info.dli_fname;
    const size_t slash = dir.find_last_of('/');
    dir = (slash == std::string::npos) ? std::string() : dir.substr(0, slash + 1);
  }
  void* h = dlopen((dir + "libaffpredict.so").c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libaffpredict.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) throw std::runtime_error(std::string("the CUDA library libaffpredict.so cannot be loaded (no CPU fallback): ") + dlerror());
#define AFF_SYM(field, name)                                                                     \
  g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(h, #name));                        \
  if (!g_api.field) throw std::runtime_error("libaffpredict.so does not export " #name);
  AFF_SYM(last_error, aff_last_error)
  AFF_SYM(scene_create, aff_scene_create)
  AFF_SYM(scene_destroy, aff_scene_destroy)
  AFF_SYM(scene_num_ik_joints, aff_scene_num_ik_joints)
  AFF_SYM(batch_create, aff_batch_create)
  AFF_SYM(batch_destroy, aff_batch_destroy)
  AFF_SYM(batch_launch, aff_batch_launch)
  AFF_SYM(batch_results, aff_batch_results)
  AFF_SYM(batch_num_steps, aff_batch_num_steps)
  AFF_SYM(batch_fetch_q, aff_batch_fetch_q)
  AFF_SYM(predict_batch, aff_predict_batch)
#undef AFF_SYM
  g_loaded = true;
  return g_api;
}

}   // namespace affgpu
