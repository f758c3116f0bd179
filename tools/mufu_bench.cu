// Developer microbenchmark: throughput of MUFU.EX2 in its fp32, f16x2 and bf16x2 forms (elements per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/mufu_bench tools/mufu_bench.cu && /tmp/mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = -0.001f * (threadIdx.x + i + 1);
    h[i] = 0xb800b800u + i;  // two small negative halves
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  const int blocks = 148, threads = 256, iters = 4096;
  cudaMalloc(&out, blocks * threads * 4);
  cudaMalloc(&cyc, blocks * 8);
  const char* names[] = {"ex2.f32", "ex2.f16x2", "ex2.bf16x2", "tanh.f32"};
  for (int m = 0; m < 4; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      if (m == 0) k<0><<<blocks, threads>>>(out, cyc, iters);
      if (m == 1) k<1><<<blocks, threads>>>(out, cyc, iters);
      if (m == 2) k<2><<<blocks, threads>>>(out, cyc, iters);
      if (m == 3) k<3><<<blocks, threads>>>(out, cyc, iters);
      cudaDeviceSynchronize();
    }
    long long c[148];
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double insts = (double)iters * 8 * threads;                 // thread-instructions per SM
    const double elems = insts * ((m == 1 || m == 2) ? 2 : 1);
    printf("MUFU_BENCH %-10s %lld cycles/SM: %.2f thread-inst/clk/SM, %.2f elements/clk/SM (%s)\n", names[m], c[0],
           insts / c[0], elems / c[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
