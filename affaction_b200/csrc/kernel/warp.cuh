// Warp primitives used by rollout.cuh.
//
// CUDA build (the product): thin wrappers over the sm_100a intrinsics.
// AFF_EMULATE build (tests/emu only, never linked into libaffpredict.so): the
// same names are provided by a fiber-based 32-lane emulator so the kernel
// source can be stepped on a CPU next to the oracle while debugging.
#ifndef AFF_KERNEL_WARP_CUH
#define AFF_KERNEL_WARP_CUH

#if defined(AFF_EMULATE)

#define AFF_DEV static inline
#define AFF_MEMBER inline
#define AFF_NOINLINE static
int w_lane();
void w_sync();
double w_shfl(double v, int src);
int w_shfl(int v, int src);
double w_shfl_xor(double v, int mask);
int w_shfl_xor(int v, int mask);
unsigned w_ballot(int pred);
template <class T> static inline T ldg(const T* p) { return *p; }
template <class T> static inline T ldg_pair(const T* p) { return *p; }
template <class T> static inline T ldg_fk(const T* p) { return *p; }
static inline double ldg_here(const double* p) { return *p; }
static inline int w_popc(unsigned v) { return __builtin_popcount(v); }
static inline int w_ffs(unsigned v) { return __builtin_ffs((int)v); }
// between the warps of a block (the emulator runs every lane of every warp as a fiber): the waits leave
// the loop for all lanes of a warp in the same scheduler round, so its barriers stay aligned
static inline double c_ld(const double* p) { return *p; }   // fibers switch only at yields
static inline void w_wait_ge(const int* flag, int v) { while (w_ballot(*(const volatile int*)flag >= v) != 0xffffffffu) {} }
static inline bool w_wait_ge_or(const int* flag, int v, const int* stop)
{
  for (;;) {
    const int s = *(const volatile int*)stop;
    if (w_ballot(*(const volatile int*)flag >= v) == 0xffffffffu) return false;
    if (w_ballot(s) == 0xffffffffu) return true;
  }
}
static inline void w_post(int* flag, int v) { w_sync(); if (w_lane() == 0) *(volatile int*)flag = v; }

#else

#define AFF_DEV __device__ __forceinline__
#define AFF_MEMBER __device__ __forceinline__
#define AFF_NOINLINE __device__ __noinline__
// opaque: kept in a register instead of being re-read from the special register at every use
AFF_DEV int w_lane() { int l = (int)(threadIdx.x & 31u); asm volatile("" : "+r"(l)); return l; }
AFF_DEV void w_sync() { __syncwarp(); }
AFF_DEV double w_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AFF_DEV int w_shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
AFF_DEV double w_shfl_xor(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
AFF_DEV int w_shfl_xor(int v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
AFF_DEV unsigned w_ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
template <class T> AFF_DEV T ldg(const T* p) { return __ldg(p); }
// A read-only load the compiler may not hoist or keep: per-lane constants that are needed a few times per
// step are re-read (L1) instead of occupying registers for the whole rollout.
AFF_DEV double ldg_here(const double* p) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
// 16-byte record through the read-only path in one LDG.128
template <class T> AFF_DEV T ldg_pair(const T* p)
{
  static_assert(sizeof(T) == 16, "ldg_pair needs a 16-byte record");
  const int4 v = __ldg(reinterpret_cast<const int4*>(p));
  T r;
  *reinterpret_cast<int4*>(&r) = v;
  return r;
}
// 32-byte record through the read-only path in two LDG.128
template <class T> AFF_DEV T ldg_fk(const T* p)
{
  static_assert(sizeof(T) == 32, "ldg_fk needs a 32-byte record");
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  const int4 b = __ldg(reinterpret_cast<const int4*>(p) + 1);
  T r;
  reinterpret_cast<int4*>(&r)[0] = a;
  reinterpret_cast<int4*>(&r)[1] = b;
  return r;
}
AFF_DEV int w_popc(unsigned v) { return __popc(v); }
AFF_DEV int w_ffs(unsigned v) { return __ffs((int)v); }
// Flags between the warps of one block (shared memory).  w_post: everything the warp wrote before is
// visible to a warp that has seen the flag.  The lanes of a waiting warp read the same word, so they
// leave the loop together.
AFF_DEV double c_ld(const double* p) { return *(const volatile double*)p; }
AFF_DEV void w_wait_ge(const int* flag, int v)
{
  while (!__all_sync(0xffffffffu, *(const volatile int*)flag >= v)) {}
  __threadfence_block();
}
AFF_DEV bool w_wait_ge_or(const int* flag, int v, const int* stop)
{
  for (;;) {
    const int s = *(const volatile int*)stop;      // read first: a posted step is served even if stop follows it
    if (__all_sync(0xffffffffu, *(const volatile int*)flag >= v)) { __threadfence_block(); return false; }
    if (__all_sync(0xffffffffu, s != 0)) return true;
  }
}
AFF_DEV void w_post(int* flag, int v)
{
  __syncwarp();
  if ((threadIdx.x & 31u) == 0u) { __threadfence_block(); *(volatile int*)flag = v; }
}

#endif

// ---- reductions built from the primitives (identical in both builds) ----

// Sum in the fixed pairwise order of the xor butterfly (offsets 16,8,4,2,1);
// the oracle's tree_sum() reproduces this order.
AFF_DEV double w_tree_sum(double v)
{
  for (int off = 16; off >= 1; off >>= 1) v = v + w_shfl_xor(v, off);
  return v;
}

AFF_DEV double w_min(double v)
{
  for (int off = 16; off >= 1; off >>= 1) { const double o = w_shfl_xor(v, off); v = (o < v) ? o : v; }
  return v;
}

AFF_DEV double w_max(double v)
{
  for (int off = 16; off >= 1; off >>= 1) { const double o = w_shfl_xor(v, off); v = (o > v) ? o : v; }
  return v;
}

// Lexicographic minimum of (value, key): smallest value, ties -> smallest key.
AFF_DEV void w_min_keyed(double& v, int& key)
{
  for (int off = 16; off >= 1; off >>= 1) {
    const double ov = w_shfl_xor(v, off);
    const int ok = w_shfl_xor(key, off);
    if (ov < v || (ov == v && ok < key)) { v = ov; key = ok; }
  }
}

// Lexicographic maximum of (value, -key): largest value, ties -> smallest key.
AFF_DEV void w_max_keyed(double& v, int& key)
{
  for (int off = 16; off >= 1; off >>= 1) {
    const double ov = w_shfl_xor(v, off);
    const int ok = w_shfl_xor(key, off);
    if (ov > v || (ov == v && ok < key)) { v = ov; key = ok; }
  }
}

#endif
