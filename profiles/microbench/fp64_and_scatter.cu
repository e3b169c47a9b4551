// Micro-benchmarks that fix the roofline denominators the fill kernel is judged against on this B200:
//   1. peak DFMA rate of the CUDA-core FP64 pipe (register-resident chains)
//   2. cost of the scatter: fp64 RED (atomicAdd, no return) vs plain store vs load+add+store, for the
//      access pattern of the element scatter (each lane writes a 3-double run at a scattered place)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_and_scatter fp64_and_scatter.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double x) {
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; k++) a[k] = threadIdx.x * 1e-3 + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = fma(a[k], x, 1e-9);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: RED, 1: plain store, 2: load+add+store.  Each thread owns `runs` runs of 3 doubles; the runs of
// the lanes of a warp are `gap` doubles apart (scattered sectors), different warps / iterations far apart.
__global__ void scatter_kernel(double *a, size_t n, int runs, int gap, int mode) {
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (int r = 0; r < runs; r++) {
    size_t base = ((t / 32) * 32 * (size_t)runs + (size_t)r * 32) * gap + (t % 32) * (size_t)gap;
    base %= (n - 4);
#pragma unroll
    for (int b = 0; b < 3; b++) {
      double v = 1.0 + b;
      if (mode == 0) atomicAdd(&a[base + b], v);
      else if (mode == 1) a[base + b] = v;
      else a[base + b] += v;
    }
  }
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float ms;
  {
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    int iters = 8192;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      dfma_kernel<<<sms * 8, 256>>>(out, iters, 1.0000001);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    double flops = 2.0 * 16 * iters * (double)sms * 8 * 256;
    printf("dfma_peak: %.2f TFLOP/s (%d SMs, %.3f ms) -> %.1f DFMA/clk/SM at %d MHz\n", flops / ms / 1e9, sms, ms,
           flops / 2 / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1000);
    cudaFree(out);
  }
  {
    size_t n = (size_t)1 << 30;  // 8 GiB of doubles, far larger than L2
    double *a;
    cudaMalloc(&a, n * sizeof(double));
    cudaMemset(a, 0, n * sizeof(double));
    const char *names[3] = {"red.f64", "plain store", "load+add+store"};
    for (int gap = 4; gap <= 64; gap *= 4) {
      for (int mode = 0; mode < 3; mode++) {
        int runs = 27, blocks = sms * 16, threads = 256;
        for (int rep = 0; rep < 2; rep++) {
          cudaEventRecord(e0);
          scatter_kernel<<<blocks, threads>>>(a, n, runs, gap, mode);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms, e0, e1);
        }
        double ops = 3.0 * runs * (double)blocks * threads;
        printf("scatter gap=%2d doubles %-15s: %.3f ms, %.2f G entries/s, %.1f entries/clk/SM\n", gap, names[mode], ms,
               ops / ms / 1e6, ops / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
      }
    }
    cudaFree(a);
  }
  return 0;
}
