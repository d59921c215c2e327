// Micro-benchmark: throughput of scattered fp64 reductions (REDG.E.ADD.F64) per SM, next to plain load+store and
// fp32 reductions, for a footprint that is L2-resident (2.5 MB per CTA) -- the yhat update of the user-space kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/red_f64 tools/micro/red_f64.cu && /tmp/red_f64
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *y, float *yf, const int *ix, int per_cta, int n_per_thread, size_t slice) {
  double *ys = y + (size_t)blockIdx.x * slice;
  float *yfs = yf + (size_t)blockIdx.x * slice;
  const int *ii = ix + (size_t)blockIdx.x * per_cta;
  for (int r = 0; r < n_per_thread; r++) {
    const int id = ii[(r * blockDim.x + threadIdx.x) % per_cta];
    if (MODE == 0) asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(ys + id), "d"(1.0) : "memory");
    if (MODE == 1) ys[id] += 1.0;
    if (MODE == 2) asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(yfs + id), "f"(1.0f) : "memory");
    if (MODE == 3) atomicAdd(ys + id, 1.0);
  }
}
int main() {
  const int ctas = 148, nt = 512, per_cta = 1 << 16, npt = 256;
  const size_t slice = 312500;
  double *y; float *yf; int *ix;
  cudaMalloc(&y, sizeof(double) * slice * ctas);
  cudaMalloc(&yf, sizeof(float) * slice * ctas);
  cudaMemset(y, 0, sizeof(double) * slice * ctas);
  cudaMemset(yf, 0, sizeof(float) * slice * ctas);
  int *h = new int[(size_t)per_cta * ctas];
  unsigned s = 12345;
  for (size_t i = 0; i < (size_t)per_cta * ctas; i++) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) % slice; }
  cudaMalloc(&ix, sizeof(int) * (size_t)per_cta * ctas);
  cudaMemcpy(ix, h, sizeof(int) * (size_t)per_cta * ctas, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char *names[4] = {"red.f64", "load+store f64", "red.f32", "atomicAdd f64 (result unused)"};
  for (int nc : {148, 16}) for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<nc, nt>>>(y, yf, ix, per_cta, npt, slice);
      if (mode == 1) k<1><<<nc, nt>>>(y, yf, ix, per_cta, npt, slice);
      if (mode == 2) k<2><<<nc, nt>>>(y, yf, ix, per_cta, npt, slice);
      if (mode == 3) k<3><<<nc, nt>>>(y, yf, ix, per_cta, npt, slice);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%3d CTAs  %-32s %8.3f ms  %.2f ns per lane-op per SM  (%.1f G ops/s total)\n", nc, names[mode], ms,
           ms * 1e6 / ((double)nt * npt), (double)nc * nt * npt / ms / 1e6);
  }
  return 0;
}
