// Micro-benchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) per SM sub-partition on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu && ./fma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float2 a[16];
  for (int i = 0; i < 16; ++i) a[i] = make_float2(seed + i, seed - i);
  const float2 w = make_float2(1.0001f, 0.9999f), b = make_float2(seed, -seed);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, w.x, b.x); a[i].y = fmaf(a[i].y, w.y, b.y); }
      else a[i] = __ffma2_rn(a[i], w, b);
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      float h;
      if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.f); else k<1><<<148, warps * 32>>>(d, iters, 1.f);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
      // FMAs per clock per SM
      double fma = 2.0 * 16 * iters * warps * 32;
      printf("%s warps/SM=%2d cycles=%.0f  fp32 FMA/clk/SM=%.1f\n", mode ? "FFMA2" : "FFMA ", warps, h, fma / h);
    }
  return 0;
}
