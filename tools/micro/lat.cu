// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/micro/lat tools/micro/lat.cu ; output of one run: profiles/r2_b200_instruction_latencies.txt
// Dependent-issue latencies of the instructions the rollout's critical path is made of (one warp, B200).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int n, double seed, int* idx)
{
  __shared__ double sh[64];
  __shared__ int shi[64];
  const int lane = threadIdx.x;
  sh[lane] = seed + lane; sh[lane + 32] = seed; shi[lane] = idx[lane]; shi[lane + 32] = idx[lane];
  __syncwarp();
  double a = seed, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = fma(a, b, c);
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // DMUL + DADD chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = a * b + c;
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  // shuffle chain (64-bit = two SHFL)
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = __shfl_sync(0xffffffffu, a, (lane + 1) & 31);
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  // LDS pointer chase
  int j = lane;
  t0 = clock64();
  for (int i = 0; i < n; ++i) j = shi[j];
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // syncwarp
  t0 = clock64();
  for (int i = 0; i < n; ++i) { __syncwarp(); asm volatile("" ::: "memory"); }
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  // STS then dependent LDS (same address)
  t0 = clock64();
  for (int i = 0; i < n; ++i) { sh[lane] = a; __syncwarp(); a = sh[(lane + 1) & 31] + 1.0; }
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // division chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = 1.0 / (a + 2.0);
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  // sqrt chain
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = sqrt(a + 2.0);
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
  // global pointer chase through L1 (ld.global.nc)
  int g = lane;
  t0 = clock64();
  for (int i = 0; i < n; ++i) g = __ldg(&idx[g]);
  t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
  // integer ALU chain
  int k = lane;
  t0 = clock64();
  for (int i = 0; i < n; ++i) k = (k * 3 + 1) ^ i;
  t1 = clock64(); if (lane == 0) cyc[9] = t1 - t0;
  // predicated select chain on doubles
  t0 = clock64();
  for (int i = 0; i < n; ++i) a = (a > 0.5) ? a - 0.25 : a + 0.75;
  t1 = clock64(); if (lane == 0) cyc[10] = t1 - t0;
  out[lane] = a + j + g + k;
}
int main()
{
  double* out; long long* cyc; int* idx;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&idx, 32 * 4);
  int h[32]; for (int i = 0; i < 32; ++i) h[i] = (i + 1) & 31;
  cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
  const int n = 4096;
  for (int rep = 0; rep < 2; ++rep) lat<<<1, 32>>>(out, cyc, n, 0.7, idx);
  long long c[16]; cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const char* nm[] = {"DFMA", "DMUL+DADD", "SHFL 64-bit", "LDS chase", "syncwarp", "STS+sync+LDS+DADD", "1/(x+2)", "sqrt(x+2)", "LDG.nc chase (L1)", "IMAD+XOR", "DSETP+select+DADD"};
  for (int i = 0; i < 11; ++i) printf("%-22s %7.1f cycles\n", nm[i], (double)c[i] / n);
  return 0;
}
