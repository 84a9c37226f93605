// The T = P F phase of the quadcopter sweep (obca_quad.cuh, kkt_solve_block) in isolation: 9 tiles x 5 DMMA k-steps, operands in shared memory,
// one barrier per repetition.  Variants: 0 = as in the kernel (loads next to their DMMA), 1 = all operand loads of a tile first,
// 2 = the (up to three) tiles of a warp interleaved, addresses hoisted out of the loop, 3 = as 2 with two accumulator chains per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/quad_tphase tools/microbench/quad_tphase.cu && build/quad_tphase
#include <cstdio>
#include <cuda_runtime.h>
constexpr int LDP = 20, LDF = 24, LDT = 24;
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int VAR>
__global__ void __launch_bounds__(128) k_t(long long* cyc, double* sink, int reps) {
  extern __shared__ double sm[];
  double *Pm = sm, *Fb = sm + 24 * LDP, *Tm = Fb + 20 * LDF, *pv = Tm + 20 * LDT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5, gid = lane >> 2, tig = lane & 3;
  for (int e = tid; e < 24 * LDP + 20 * LDF + 20 * LDT + 24; e += blockDim.x) sm[e] = 1e-3 * (e % 17);
  __syncthreads();
  const long long t0 = clock64();
  // variants 2, 3: tile plan of this warp (nwarp >= 3: one chunk of up to three tiles)
  int tA[3], tB[3], tS[3], tP[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int idx0 = warp + j * nwarp; const bool on = idx0 < 9; const int idx = on ? idx0 : warp;
    const int m = idx / 3, n = idx % 3, row = m * 8 + gid;
    tA[j] = row * LDP + tig; tB[j] = tig * LDF + n * 8 + gid;
    tS[j] = (on && row < 20) ? row * LDT + n * 8 + tig * 2 : -1;
    tP[j] = (on && n == 2 && tig == 2) ? row : -1;
  }
  for (int r = 0; r < reps; ++r) {
    if (VAR >= 2) {
      double c[3][2], e[3][2];
#pragma unroll
      for (int j = 0; j < 3; ++j) { c[j][0] = c[j][1] = e[j][0] = e[j][1] = 0.0; }
      if (VAR == 2) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks)
#pragma unroll
          for (int j = 0; j < 3; ++j) dmma(c[j][0], c[j][1], Pm[tA[j] + ks * 4], Fb[tB[j] + ks * 4 * LDF]);
      } else {
#pragma unroll
        for (int ks = 0; ks < 3; ++ks)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            dmma(c[j][0], c[j][1], Pm[tA[j] + ks * 4], Fb[tB[j] + ks * 4 * LDF]);
            if (ks < 2) dmma(e[j][0], e[j][1], Pm[tA[j] + (ks + 3) * 4], Fb[tB[j] + (ks + 3) * 4 * LDF]);
          }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        c[j][0] += e[j][0]; c[j][1] += e[j][1];
        if (tP[j] >= 0) c[j][1] += pv[tP[j]];
        if (tS[j] >= 0) *reinterpret_cast<double2*>(Tm + tS[j]) = make_double2(c[j][0], c[j][1]);
      }
    } else
    for (int idx = warp; idx < 9; idx += nwarp) {
      const int m = idx / 3, n = idx % 3;
      double c0 = 0.0, c1 = 0.0;
      const double* const ap = Pm + (m * 8 + gid) * LDP + tig;
      const double* const bp = Fb + tig * LDF + n * 8 + gid;
      if (VAR == 0) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) dmma(c0, c1, ap[ks * 4], bp[ks * 4 * LDF]);
      } else {
        double a[5], b[5];
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) { a[ks] = ap[ks * 4]; b[ks] = bp[ks * 4 * LDF]; }
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) dmma(c0, c1, a[ks], b[ks]);
      }
      const int row = m * 8 + gid;
      if (row < 20) {
        if (n == 2 && tig == 2) c1 += pv[row];
        *reinterpret_cast<double2*>(Tm + row * LDT + n * 8 + tig * 2) = make_double2(c0, c1);
      }
    }
    __syncthreads();
    if (tid < 24) Pm[tid * LDP + (r % 17)] = Tm[(r % 17) * LDT + tid] * 1e-3;      // a dependence between repetitions
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0) { cyc[blockIdx.x] = t1 - t0; sink[blockIdx.x] = Tm[5]; }
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* cyc; double* sink;
  cudaMalloc(&cyc, 8 * sms * 8); cudaMalloc(&sink, 8 * sms * 8);
  const int reps = 2000; const size_t smem = (24 * LDP + 20 * LDF + 20 * LDT + 24) * 8;
  printf("{");
  for (int var = 0; var < 4; ++var)
    for (int per_sm : {1, 2, 4}) {
      const int grid = sms * per_sm;
      for (int it = 0; it < 2; ++it) {
        if (var == 0) k_t<0><<<grid, 128, smem>>>(cyc, sink, reps); else if (var == 1) k_t<1><<<grid, 128, smem>>>(cyc, sink, reps);
        else if (var == 2) k_t<2><<<grid, 128, smem>>>(cyc, sink, reps); else k_t<3><<<grid, 128, smem>>>(cyc, sink, reps);
        cudaDeviceSynchronize();
      }
      long long* h = new long long[grid]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < grid; ++i) s += h[i]; delete[] h;
      printf("\"T_phase_var%d_%dcta_per_sm_cycles\": %.1f, ", var, per_sm, s / grid / reps);
    }
  printf("\"note\": \"cycles per repetition = one T phase (w0: 3 tiles = 15 DMMA) + 2 barriers\"}\n");
  return 0;
}
