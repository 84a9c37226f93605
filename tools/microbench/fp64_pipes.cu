// FP64 vector (DFMA) and tensor (DMMA.8x8x4) latency / throughput of the device, measured with clock64 inside the kernel.
// Development tool behind DESIGN.md section 5 ("what bounds the latency chains").  Build + run (on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_pipes tools/microbench/fp64_pipes.cu && build/fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// CH independent dependent-chains of DFMA per thread, ITER steps each
template <int CH>
__global__ void k_dfma(double* out, long long* cyc, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = threadIdx.x * 1e-3 + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH>
__global__ void k_dmma(double* out, long long* cyc, int iters, double a, double b) {
  double x[CH][2];
#pragma unroll
  for (int c = 0; c < CH; ++c) { x[c][0] = threadIdx.x * 1e-3 + c; x[c][1] = 1.0; }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) dmma(x[c][0], x[c][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c][0] + x[c][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static double run(F launch, int grid, long long* dcyc) {
  launch();
  cudaDeviceSynchronize();
  launch();
  cudaDeviceSynchronize();
  long long* h = new long long[grid];
  cudaMemcpy(h, dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0.0;
  for (int i = 0; i < grid; ++i) s += (double)h[i];
  delete[] h;
  return s / grid;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, iters = 4096;
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms * 16);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d,\n", p.name, sms, p.clockRate / 1000);
  // latency: one warp on one SM, one chain
  double c = run([&] { k_dfma<1><<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, 1, cyc);
  printf(" \"dfma_latency_cycles\": %.2f,\n", c / iters);
  c = run([&] { k_dmma<1><<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, 1, cyc);
  printf(" \"dmma_latency_cycles\": %.2f,\n", c / iters);
  // one warp, 8 independent chains: issue interval of a single warp
  c = run([&] { k_dfma<8><<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, 1, cyc);
  printf(" \"dfma_one_warp_8chains_cycles_per_inst\": %.2f,\n", c / (iters * 8.0));
  c = run([&] { k_dmma<8><<<1, 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, 1, cyc);
  printf(" \"dmma_one_warp_8chains_cycles_per_inst\": %.2f,\n", c / (iters * 8.0));
  // throughput: every SM full of warps (1024 threads per CTA, one CTA per SM), 8 chains per thread
  for (int warps : {4, 8, 16, 32}) {
    c = run([&] { k_dfma<8><<<sms, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, sms, cyc);
    printf(" \"dfma_fma_per_clk_per_sm_%dwarps\": %.2f,\n", warps, (double)iters * 8 * warps * 32 / c);
    c = run([&] { k_dmma<8><<<sms, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, sms, cyc);
    printf(" \"dmma_fma_per_clk_per_sm_%dwarps\": %.2f,\n", warps, (double)iters * 8 * warps * 256 / c);
  }
  // dependent single chains with many warps: how latency-bound code scales with resident warps
  for (int warps : {4, 8, 12, 16}) {
    c = run([&] { k_dfma<1><<<sms, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, sms, cyc);
    printf(" \"dfma_dependent_cycles_per_inst_%dwarps\": %.2f,\n", warps, c / iters);
    c = run([&] { k_dmma<1><<<sms, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9); }, sms, cyc);
    printf(" \"dmma_dependent_cycles_per_inst_%dwarps\": %.2f,\n", warps, c / iters);
  }
  printf(" \"note\": \"cycles = SM clock (clock64); fma counts: DFMA = 32 per warp instruction, DMMA.8x8x4 = 256\"}\n");
  return 0;
}
