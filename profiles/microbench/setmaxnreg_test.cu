// Does warpgroup register re-allocation (setmaxnreg) work the way fill_kernel wants to use it on sm_100a?
// 256 threads, 2 CTAs/SM at 128 regs; per iteration warpgroup 0 grows to HI, warpgroup 1 shrinks to LO.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int N> __device__ __forceinline__ void inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int HI, int LO, bool PRE>
__global__ void __launch_bounds__(256, 2) k(double *out, const double *in, int iters) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  double acc = tid;
  for (int it = 0; it < iters; it++) {
    if (PRE) {  // a phase that uses many registers in every warp before the split
      double t[40];
#pragma unroll
      for (int q = 0; q < 40; q++) t[q] = in[(tid * 40 + q + it) & 1023];
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 40; q++) acc += t[q] * t[(q + 7) % 40];
      sm[tid] = acc;
      __syncthreads();
      acc += sm[(tid + 1) & 255];
    }
    if (tid < 128) {
      inc<HI>();
      double a[60];
#pragma unroll
      for (int q = 0; q < 60; q++) a[q] = acc + q;
#pragma unroll 1
      for (int g = 0; g < 27; g++)
#pragma unroll
        for (int q = 0; q < 60; q++) a[q] = fma(a[q], 1.0000001, sm[g]);
#pragma unroll
      for (int q = 0; q < 60; q++) acc += a[q];
      dec<128>();
    } else {
      dec<LO>();
      acc += sm[tid];
      inc<128>();
    }
    __syncthreads();
  }
  out[blockIdx.x * 256 + tid] = acc;
}
template <int HI, int LO, bool PRE>
void run(const char *name, double *out, double *in, int grid, int smem) {
  cudaFuncSetAttribute(k<HI, LO, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<HI, LO, PRE><<<grid, 256, smem>>>(out, in, 10);
  printf("%s grid %d smem %d: %s\n", name, grid, smem, cudaGetErrorString(cudaDeviceSynchronize()));
  fflush(stdout);
}
int main(int argc, char **argv) {
  double *out, *in;
  cudaMalloc(&out, 8 * 256 * 1024);
  cudaMalloc(&in, 8 * 1024);
  cudaMemset(in, 0, 8 * 1024);
  int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0) run<216, 40, false>("216/40", out, in, 8, 2048);
  if (which == 1) run<216, 40, false>("216/40", out, in, 8, 107 * 1024);
  if (which == 2) run<216, 40, true>("216/40 pre", out, in, 8, 107 * 1024);
  if (which == 3) run<208, 48, true>("208/48 pre", out, in, 8, 107 * 1024);
  if (which == 4) run<216, 40, true>("216/40 pre", out, in, 296, 107 * 1024);
  return 0;
}
