// Micro-benchmark: FP64 tensor-core (DMMA, mma.sync ... f64) issue rate on sm_100a against the DFMA pipe.
// Decides whether the element-block contraction of the fill kernel should go through mma.sync (north_star:
// "FP64 DMMA tensor cores ... only if ncu shows they beat the CUDA-core path").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void dmma_kernel(double *out, int iters) {
  double c[NACC][4];
  for (int t = 0; t < NACC; t++)
    for (int k = 0; k < 4; k++) c[t][k] = 0.0;
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, b0 = 0.5, b1 = 0.25;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int t = 0; t < NACC; t++) {
      if (SHAPE == 884)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a0), "d"(b0));
      else if (SHAPE == 1684)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3]) : "d"(a0), "d"(a1), "d"(b0));
      else if (SHAPE == 1688)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                     : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    }
  }
  double s = 0.0;
  for (int t = 0; t < NACC; t++)
    for (int k = 0; k < 4; k++) s += c[t][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_kernel(double *out, int iters) {
  double c[16];
  for (int t = 0; t < 16; t++) c[t] = t;
  double a = threadIdx.x * 1e-3, b = 0.999;
  for (int it = 0; it < iters; it++)
#pragma unroll
    for (int t = 0; t < 16; t++) c[t] = fma(c[t], b, a);
  double s = 0.0;
  for (int t = 0; t < 16; t++) s += c[t];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved in one instruction stream (8 + 16 independent accumulators): if the tensor sub-pipe and
// the FP64 pipe were separate units the sum would exceed either peak
__global__ void mixed_kernel(double *out, int iters) {
  double c[8][2], d[16];
  for (int t = 0; t < 8; t++) c[t][0] = c[t][1] = 0.0;
  for (int t = 0; t < 16; t++) d[t] = t;
  double a0 = threadIdx.x * 1e-3, b0 = 0.5, b = 0.999;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int t = 0; t < 8; t++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a0), "d"(b0));
      d[2 * t] = fma(d[2 * t], b, a0);
      d[2 * t + 1] = fma(d[2 * t + 1], b, a0);
    }
  }
  double s = 0.0;
  for (int t = 0; t < 8; t++) s += c[t][0] + c[t][1];
  for (int t = 0; t < 16; t++) s += d[t];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int threads = 32 * warps, grid = sms;
    float ms;
    ms = time_ms([&] { dfma_kernel<<<grid, threads>>>(out, iters); });
    printf("warps/SM %2d  DFMA           %7.2f TFLOP/s\n", warps, 2.0 * 16 * iters * (double)threads * grid / ms / 1e9);
    ms = time_ms([&] { dmma_kernel<884, 8><<<grid, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA m8n8k4    %7.2f TFLOP/s\n", warps, 2.0 * 256 * 8 * iters * (double)warps * grid / ms / 1e9);
    ms = time_ms([&] { dmma_kernel<1684, 8><<<grid, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA m16n8k4   %7.2f TFLOP/s\n", warps, 2.0 * 512 * 8 * iters * (double)warps * grid / ms / 1e9);
    ms = time_ms([&] { mixed_kernel<<<grid, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA + DFMA    %7.2f TFLOP/s  (DMMA part %.2f, DFMA part %.2f)\n", warps,
           (2.0 * 256 * 8 * warps + 2.0 * 16 * threads) * iters * (double)grid / ms / 1e9,
           2.0 * 256 * 8 * iters * (double)warps * grid / ms / 1e9, 2.0 * 16 * iters * (double)threads * grid / ms / 1e9);
    ms = time_ms([&] { dmma_kernel<1688, 8><<<grid, threads>>>(out, iters); });
    printf("warps/SM %2d  DMMA m16n8k8   %7.2f TFLOP/s\n", warps, 2.0 * 1024 * 8 * iters * (double)warps * grid / ms / 1e9);
  }
  return 0;
}
