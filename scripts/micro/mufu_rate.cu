// Micro-benchmark: MUFU.TANH / MUFU.EX2 throughput per SM, alone and mixed with FFMA2, on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: tanh only; 1: ex2 only; 2: the swish form (FFMA2, 2 x tanh, FFMA2) per element pair
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + 0.01f * i + threadIdx.x * 1e-3f, seed - 0.01f * i);
  const float2 half = make_float2(0.5f, 0.5f), b = make_float2(1e-3f, -1e-3f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i].x));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i].y));
      } else if (MODE == 1) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].x));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i].y));
      } else {
        const float2 h = __ffma2_rn(a[i], half, b);
        float2 t;
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
        a[i] = __ffma2_rn(h, t, h);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}

int main() {
  float* d; cudaMalloc(&d, 1 << 24);
  const int iters = 4096;
  const char* names[3] = {"tanh ", "ex2  ", "swish"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      float h;
      if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 0.3f);
      else if (mode == 1) k<1><<<148, warps * 32>>>(d, iters, 0.3f);
      else k<2><<<148, warps * 32>>>(d, iters, 0.3f);
      cudaDeviceSynchronize();
      cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
      double ops = 2.0 * 8 * iters * warps * 32;     // transcendental results
      printf("%s warps/SM=%2d cycles=%.0f  results/clk/SM=%.2f\n", names[mode], warps, h, ops / h);
    }
  return 0;
}
