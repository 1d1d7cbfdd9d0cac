// dpx_rates.cu -- issue rates of the instructions the banded-DP kernels are built from, on the device it runs on.
// Measurement tool (profiles/): dependency-free chains, 8 per thread, 148 x 8 CTAs of 256 threads.
//   mode 0: VIADDMNMX (s32)             mode 1: VIADDMNMX.S16x2
//   mode 2: the packed row step's cell pair: IMAD.SHL + LOP3 + 2 x VIADDMNMX.S16x2
//   mode 3: the 32-bit row step's cell: 2 x VIADDMNMX (+ IMAD.MOV)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_rates dpx_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned* out, int iters, unsigned a0, unsigned b0, unsigned m0) {
  unsigned x[8], y[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    x[c] = a0 + threadIdx.x * (c + 1);
    y[c] = b0 + c;
  }
  unsigned m = m0 + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (MODE == 0) {
          x[c] = (unsigned)__viaddmax_s32((int)x[c], (int)b0, (int)y[c]);
        } else if (MODE == 1) {
          x[c] = __viaddmax_s16x2(x[c], b0, y[c]);
        } else if (MODE == 2) {
          const unsigned v = m << (r * 8 + c & 15);
          const unsigned a2 = b0 | (~v & 0x80008000u);
          y[c] = __viaddmax_s16x2(x[c], a2, y[c]);
          x[c] = __viaddmax_s16x2(y[c], a0, x[c]);
        } else {
          y[c] = (unsigned)__viaddmax_s32((int)x[c], (int)b0, (int)y[c]);
          x[c] = (unsigned)__viaddmax_s32((int)y[c], (int)a0, (int)x[c]);
        }
      }
    }
    m = m * 3 + 1;
  }
  unsigned s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += x[c] ^ y[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static double run(unsigned* out, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * 8, block = 256;
  k<MODE><<<grid, block>>>(out, iters / 10, 3, 1000, 1);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(out, iters, 3, 1000, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  const double groups = (double)iters * 4 * 8 * grid * block;  // (chain, round) steps executed by all threads
  return groups / (best * 1e-3);
}

int main() {
  unsigned* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 4000;
  const double r0 = run<0>(out, iters), r1 = run<1>(out, iters), r2 = run<2>(out, iters), r3 = run<3>(out, iters);
  printf("{\"viaddmnmx_s32_Tips\": %.3f, \"viaddmnmx_s16x2_Tips\": %.3f, "
         "\"packed_cell_pairs_Tps\": %.3f, \"packed_cells_Tps\": %.3f, \"s32_cells_Tps\": %.3f}\n",
         r0 / 1e12, r1 / 1e12, r2 / 1e12, 2 * r2 / 1e12, r3 / 1e12);
  return cudaGetLastError() != cudaSuccess;
}
