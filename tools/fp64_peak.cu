// Microbenchmark: FP64 throughput of the DMMA (mma.sync m8n8k4.f64) and DFMA pipes on the box's GPU.
// These are the denominators context for the posterior GEMM roofline (MEASURED_PEAKS.json has no fp64 figure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu && tools/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0000001;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int kind = 0; kind < 2; ++kind) {
      int iters = 20000;
      dim3 grid(sms * 2), block(warps * 32 / 2);
      if (kind == 0) dmma_kernel<<<grid, block>>>(out, 100); else dfma_kernel<<<grid, block>>>(out, 100);
      cudaDeviceSynchronize();
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) dmma_kernel<<<grid, block>>>(out, iters); else dfma_kernel<<<grid, block>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      double threads = (double)grid.x * block.x;
      double flops = kind == 0 ? (threads / 32.0) * iters * 16.0 * 512.0 : threads * iters * 16.0 * 2.0;
      printf("%s warps/SM=%d : %.2f TFLOP/s (%.3f ms)\n", kind == 0 ? "DMMA.8x8x4" : "DFMA      ", warps, flops / best * 1e-9, best);
    }
  }
  printf("device: %s, %d SMs, clock %d MHz\n", p.name, sms, p.clockRate / 1000);
  return 0;
}
