// Microbenchmarks that size the channelwise/stem kernels: fp32 FMA issue rate with FFMA vs the
// packed FFMA2, bf16 HFMA2, and a plain 128-bit copy.   nvcc -arch=sm_100a -O3 -o microbench microbench.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void fma_kernel(float* out, int iters, float a0) {
  float2 acc[8];
  __nv_bfloat162 hacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i); hacc[i] = __floats2bfloat162_rn(i, i + 1); }
  float2 a = make_float2(a0, a0 * 0.5f), b = make_float2(0.25f, 0.125f);
  __nv_bfloat162 ha = __floats2bfloat162_rn(a0, a0), hb = __floats2bfloat162_rn(0.25f, 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
      if (MODE == 1) { acc[i] = __ffma2_rn(acc[i], a, b); }
      if (MODE == 2) { hacc[i] = __hfma2(hacc[i], ha, hb); }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y + __low2float(hacc[i]) + __high2float(hacc[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 with the operand pattern of the channelwise stencil: acc[set][col] += x[jj] * w[set][dw].
// MODE 0: 9 consecutive FFMA2 share x (operand reuse possible); MODE 1: x, w and acc all change
// from one FFMA2 to the next (3 distinct register pairs per instruction, no reuse).
template <int MODE>
__global__ void __launch_bounds__(256, 2) stencil_fma_kernel(const float2* __restrict__ src, float* out, int iters) {
  float2 x[10], w[9], acc[3][8];
#pragma unroll
  for (int i = 0; i < 10; ++i) x[i] = src[threadIdx.x + 32 * i];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = src[threadIdx.x + 32 * (10 + i)];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[a][j] = make_float2(a, j);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int jj = 0; jj < 10; ++jj)
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          const int j = jj - dw;
          if (j >= 0 && j < 8) {
#pragma unroll
            for (int a = 0; a < 3; ++a) acc[a][j] = __ffma2_rn(x[jj], w[a * 3 + dw], acc[a][j]);
          }
        }
    } else {
#pragma unroll
      for (int i = 0; i < 72; ++i)
        acc[i % 3][(i / 3) % 8] = __ffma2_rn(x[i % 10], w[i % 9], acc[i % 3][(i / 3) % 8]);
    }
  }
  float s = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += acc[a][j].x + acc[a][j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Issue-slot test: 8 independent FFMA2 chains interleaved with NALU integer (ALU-pipe) ops per
// 8 FFMA2.  If an FFMA2 only held the FMA pipe for its second cycle, ALU work up to one op per
// FFMA2 would be free; if it also holds the issue port, time grows as 2*FFMA2 + ALU.
template <int NALU>
__global__ void __launch_bounds__(256, 2) mix_kernel(float* out, int iters, float a0, unsigned k0) {
  float2 acc[8];
  unsigned v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i); v[i] = threadIdx.x * 2654435761u + i; }
  const float2 a = make_float2(a0, a0 * 0.5f), b = make_float2(0.25f, 0.125f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] = __ffma2_rn(acc[i], a, b);
      if (i < NALU) v[i] = (v[i] ^ k0) & 0xfffffff7u;          // one LOP3 per op, independent chains
      if (i + 8 < NALU) v[i] = (v[i] << 16) | (v[i] >> 27);    // SHF
    }
  }
  float s = 0;
  unsigned u = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { s += acc[i].x + acc[i].y; u ^= v[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(u & 1);
}

__global__ void copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) out[i] = in[i];
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = p.multiProcessorCount * 4, threads = 512;
  const char* names[3] = {"FFMA (2 per lane-iter)", "FFMA2", "HFMA2.BF16"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) fma_kernel<0><<<blocks, threads>>>(out, iters, 1.0001f);
      if (mode == 1) fma_kernel<1><<<blocks, threads>>>(out, iters, 1.0001f);
      if (mode == 2) fma_kernel<2><<<blocks, threads>>>(out, iters, 1.0001f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fmas = (double)blocks * threads * iters * 8 * 2;
      if (rep) printf("%-24s %8.3f ms  %8.2f TFMA/s  (%.1f FMA/clk/SM at %d MHz nominal)\n", names[mode], ms,
                      fmas / ms / 1e9, fmas / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
    }
  }
  {
    float2* src; cudaMalloc(&src, 32 * 32 * sizeof(float2)); cudaMemset(src, 0, 32 * 32 * sizeof(float2));
    const char* nm[2] = {"FFMA2 stencil pattern (x shared by 9)", "FFMA2 3 distinct operands, no reuse"};
    for (int mode = 0; mode < 2; ++mode)
      for (int rep = 0; rep < 2; ++rep) {
        const int it2 = 4000, bl = p.multiProcessorCount * 2, th = 256;
        cudaEventRecord(e0);
        if (mode == 0) stencil_fma_kernel<0><<<bl, th>>>(src, out, it2);
        else stencil_fma_kernel<1><<<bl, th>>>(src, out, it2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fmas = (double)bl * th * it2 * 72 * 2;
        if (rep) printf("%-40s %8.3f ms  %8.2f TFMA/s  (%.1f FMA/clk/SM at %d MHz nominal)\n", nm[mode], ms,
                        fmas / ms / 1e9, fmas / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
      }
  }
  {
    const int it3 = 20000, bl = p.multiProcessorCount * 2, th = 256;   // 16 warps per SM
    for (int nalu = 0; nalu <= 16; nalu += 4)
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (nalu == 0) mix_kernel<0><<<bl, th>>>(out, it3, 1.0001f, 0x5a5a5a5au);
        if (nalu == 4) mix_kernel<4><<<bl, th>>>(out, it3, 1.0001f, 0x5a5a5a5au);
        if (nalu == 8) mix_kernel<8><<<bl, th>>>(out, it3, 1.0001f, 0x5a5a5a5au);
        if (nalu == 12) mix_kernel<12><<<bl, th>>>(out, it3, 1.0001f, 0x5a5a5a5au);
        if (nalu == 16) mix_kernel<16><<<bl, th>>>(out, it3, 1.0001f, 0x5a5a5a5au);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double cyc = ms * 1e-3 * p.clockRate * 1e3;               // SM cycles at nominal clock
        const double warp_ffma2_per_smsp = (double)bl * th / 32 * it3 * 8 / p.multiProcessorCount / 4;
        if (rep) printf("8 FFMA2 + %2d ALU ops per iteration: %8.3f ms  %.2f cycles per FFMA2 per scheduler\n", nalu, ms,
                        cyc / warp_ffma2_per_smsp);
      }
  }
  size_t n = (size_t)1 << 26;   // 1 GiB in uint4
  uint4 *a, *b; cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMemset(a, 1, n * 16);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    copy_kernel<<<p.multiProcessorCount * 8, 512>>>(a, b, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep) printf("copy 1 GiB: %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * n * 16 / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
