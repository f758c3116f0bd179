// Developer check of the FMA-pipe exp2 emulation (common.cuh: ex2_emu2) against exp2f on the device.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I infinicube_b200/csrc -o /tmp/emu_test tools/emu_test.cu && /tmp/emu_test
#include <cstdio>
#include <cmath>
#include <vector>
#include "common.cuh"
using namespace icb;

__global__ void k(const float* x, float* y_emu, float* y_mufu, int n, float magic_rt) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 >= n) return;
  unsigned long long r = ex2_emu2(pk2(x[i], x[i + 1]));
  float a, b;
  upk2(r, a, b);
  y_emu[i] = a;
  y_emu[i + 1] = b;
  y_mufu[i] = ex2_approx(x[i]);
  y_mufu[i + 1] = ex2_approx(x[i + 1]);
}

int main() {
  const int n = 1 << 20;
  std::vector<float> x(n), ye(n), ym(n);
  for (int i = 0; i < n; ++i) x[i] = -130.0f + 138.0f * (float)i / n;
  float *dx, *de, *dm;
  cudaMalloc(&dx, n * 4);
  cudaMalloc(&de, n * 4);
  cudaMalloc(&dm, n * 4);
  cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice);
  k<<<n / 2 / 256, 256>>>(dx, de, dm, n, 12582912.0f);
  cudaMemcpy(ye.data(), de, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ym.data(), dm, n * 4, cudaMemcpyDeviceToHost);
  double worst_e[2] = {0, 0}, worst_m = 0;
  int arg_e[2] = {-1, -1};
  for (int i = 0; i < n; ++i) {
    if (x[i] < -125.9f) continue;
    const double ref = exp2((double)x[i]);
    const double ee = fabs(ye[i] - ref) / ref, em = fabs(ym[i] - ref) / ref;
    if (ee > worst_e[i & 1]) worst_e[i & 1] = ee, arg_e[i & 1] = i;
    if (em > worst_m) worst_m = em;
  }
  printf("EMU_TEST lo-lane max rel err %.3e at x=%g (emu %g), hi-lane %.3e at x=%g (emu %g), mufu %.3e, cuda: %s\n",
         worst_e[0], arg_e[0] >= 0 ? x[arg_e[0]] : 0.f, arg_e[0] >= 0 ? ye[arg_e[0]] : 0.f, worst_e[1],
         arg_e[1] >= 0 ? x[arg_e[1]] : 0.f, arg_e[1] >= 0 ? ye[arg_e[1]] : 0.f, worst_m,
         cudaGetErrorString(cudaGetLastError()));
  for (float t : {0.0f, -0.3f, -1.0f, -1.5f, -7.25f, 3.2f}) {
    int i = (int)((t + 130.0f) / 138.0f * n);
    printf("  x=%g emu=%g mufu=%g | x=%g emu=%g mufu=%g\n", x[i & ~1], ye[i & ~1], ym[i & ~1], x[i | 1], ye[i | 1], ym[i | 1]);
  }
  return 0;
}
