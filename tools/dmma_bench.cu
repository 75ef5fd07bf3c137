// FP64 pipe microbenchmark for sm_100a: DFMA (CUDA cores) vs DMMA (mma.sync.m8n8k4.f64 tensor path) vs both at once.
// north_star: "FP64 DMMA tensor cores only where ncu shows the small contractions actually beat CUDA-core FMA".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gpurun_out/dmma_bench tools/dmma_bench.cu && gpurun_out/dmma_bench
// Prints TFLOP/s of: (A) register-resident DFMA chains, (B) DMMA chains, (C) DFMA and DMMA interleaved in every warp,
// (D) half of the warps DFMA / half DMMA, (E) an 8x8 line contraction (the n = 8 tensor stage: out[a][l] = sum_m D[a][m] u[m][l])
// from shared memory to shared memory with DFMA (one thread per line, table in registers) and (F) the same with DMMA.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE> __global__ void __launch_bounds__(256) peak_kernel(double *out, int iters, double a, double b)
{
  double x[8], c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = threadIdx.x + i; c[i][0] = i; c[i][1] = -i; }
  const bool warp_dfma = MODE == 3 ? ((threadIdx.x >> 5) & 1) == 0 : true;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      if (MODE == 0 || MODE == 2 || (MODE == 3 && warp_dfma)) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
      }
      if (MODE == 1 || MODE == 2 || (MODE == 3 && !warp_dfma)) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i] + c[i][0] + c[i][1];
  out[(long)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (E) one thread per line of 8: load 8, 64 DFMA with the table in registers (constant bank), store 8.  256 threads = 4 elements of 8^3 (64 lines each)
struct Tab8 { double D[64]; };
__global__ void __launch_bounds__(256) line_dfma_kernel(const __grid_constant__ Tab8 T, double *out, int iters)
{
  __shared__ double A[256 * 9], G[256 * 9];
  const int t = threadIdx.x;
  for (int i = 0; i < 8; i++) A[t * 9 + i] = t + i;
  __syncthreads();
  for (int it = 0; it < iters; it++) {
    double r[8];
#pragma unroll
    for (int m = 0; m < 8; m++) r[m] = A[t * 9 + m];
#pragma unroll
    for (int a = 0; a < 8; a++) {
      double s = 0;
#pragma unroll
      for (int m = 0; m < 8; m++) s = fma(T.D[a * 8 + m], r[m], s);
      G[t * 9 + a] = s;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 8; m++) r[m] = G[t * 9 + m];
#pragma unroll
    for (int a = 0; a < 8; a++) {
      double s = 0;
#pragma unroll
      for (int m = 0; m < 8; m++) s = fma(T.D[a * 8 + m], r[m], s);
      A[t * 9 + a] = s * 1e-3;
    }
    __syncthreads();
  }
  out[(long)blockIdx.x * 256 + t] = A[t * 9];
}
// (F) the same contraction with DMMA: a warp owns 32 lines = 4 groups of 8; per group C[8 outputs][8 lines] = D[8x8] U[8 x 8 lines] in two k = 4 steps.
// Fragments (PTX m8n8k4 .f64): A row = lane/4, col = lane%4 ; B row(k) = lane%4, col(n) = lane/4 ; C row = lane/4, cols = 2*(lane%4)+{0,1}.
__global__ void __launch_bounds__(256) line_dmma_kernel(const __grid_constant__ Tab8 T, double *out, int iters)
{
  __shared__ double A[256 * 9], G[256 * 9];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  for (int i = 0; i < 8; i++) A[t * 9 + i] = t + i;
  const double a0 = T.D[(lane >> 2) * 8 + (lane & 3)], a1 = T.D[(lane >> 2) * 8 + 4 + (lane & 3)];
  __syncthreads();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int line = w * 32 + g * 8;
      const double b0 = A[(line + (lane >> 2)) * 9 + (lane & 3)], b1 = A[(line + (lane >> 2)) * 9 + 4 + (lane & 3)];
      double c0 = 0, c1 = 0;
      dmma(c0, c1, a0, b0);
      dmma(c0, c1, a1, b1);
      G[(line + 2 * (lane & 3)) * 9 + (lane >> 2)] = c0;
      G[(line + 2 * (lane & 3) + 1) * 9 + (lane >> 2)] = c1;
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 4; g++) {
      const int line = w * 32 + g * 8;
      const double b0 = G[(line + (lane >> 2)) * 9 + (lane & 3)], b1 = G[(line + (lane >> 2)) * 9 + 4 + (lane & 3)];
      double c0 = 0, c1 = 0;
      dmma(c0, c1, a0, b0);
      dmma(c0, c1, a1, b1);
      A[(line + 2 * (lane & 3)) * 9 + (lane >> 2)] = c0 * 1e-3;
      A[(line + 2 * (lane & 3) + 1) * 9 + (lane >> 2)] = c1 * 1e-3;
    }
    __syncthreads();
  }
  out[(long)blockIdx.x * 256 + t] = A[t * 9];
}

template <class F> static float time_ms(F f)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int blocks = prop.multiProcessorCount * 8, iters = 2048;
  double *out;
  CK(cudaMalloc(&out, (size_t)blocks * 256 * sizeof(double)));
  printf("{\"device\": \"%s\", \"sms\": %d", prop.name, prop.multiProcessorCount);
  const double per_thread_dfma = 2.0 * 8 * 8 * iters, per_warp_dmma = 512.0 * 8 * 8 * iters;
  float ms;
  ms = time_ms([&] { peak_kernel<0><<<blocks, 256>>>(out, iters, 0.999999, 1e-9); });
  printf(", \"dfma_tflops\": %.2f", per_thread_dfma * blocks * 256 / ms / 1e9);
  ms = time_ms([&] { peak_kernel<1><<<blocks, 256>>>(out, iters, 0.999999, 1e-9); });
  printf(", \"dmma_tflops\": %.2f", per_warp_dmma * blocks * 8 / ms / 1e9);
  ms = time_ms([&] { peak_kernel<2><<<blocks, 256>>>(out, iters, 0.999999, 1e-9); });
  printf(", \"interleaved_in_warp_tflops\": %.2f", (per_thread_dfma * blocks * 256 + per_warp_dmma * blocks * 8) / ms / 1e9);
  ms = time_ms([&] { peak_kernel<3><<<blocks, 256>>>(out, iters, 0.999999, 1e-9); });
  printf(", \"split_warps_tflops\": %.2f", (per_thread_dfma * blocks * 128 + per_warp_dmma * blocks * 4) / ms / 1e9);
  Tab8 T;
  for (int i = 0; i < 64; i++) T.D[i] = 0.01 * ((i * 37) % 11 - 5);
  const int it2 = 512;
  const double fl = 2.0 * 2 * 64 * 256 * (double)it2 * blocks;       // two passes x 64 FMA per line x 256 lines
  ms = time_ms([&] { line_dfma_kernel<<<blocks, 256>>>(T, out, it2); });
  printf(", \"line8_dfma_tflops\": %.2f", fl / ms / 1e9);
  ms = time_ms([&] { line_dmma_kernel<<<blocks, 256>>>(T, out, it2); });
  printf(", \"line8_dmma_tflops\": %.2f}\n", fl / ms / 1e9);
  CK(cudaGetLastError());
  return 0;
}
