// Microbenchmark of the planar channelwise stencil's inner loop (csrc/x3d_dw_planar.cu): packed FFMA2 whose
// tap operand is a UNIFORM register, fed by LDS.64, with and without the epilogue's work.
//   nvcc -arch=sm_100a -O3 -o microbench_ur microbench_ur.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

__constant__ float2 c_w[8 * 28];

__device__ __forceinline__ float2 lds2(unsigned a) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}

// MODE 0: x from registers (pure FFMA2 R,R,UR,R stream)   1: x by LDS.64 (30 per 216 FFMA2)
// MODE 2: + per step the epilogue of one output frame (8 x: bias add, swish, bf16 pack, 4-byte store)
// MODE 3: as 1 but the taps in ordinary registers (R,R,R,R form)
// MODE 4: as 1 + all 27 taps re-read from the constant bank every step (27 LDCU.64 per 216 FFMA2)
// MODE 5: as 1 + the epilogue WITHOUT swish (bias add, pack, store, SE sum), accumulator set chosen statically
// MODE 6: as 2 with the accumulator set chosen statically (steps unrolled by 3)
template <int MODE, int Q>
__global__ void __launch_bounds__(256, MODE == 3 ? 2 : 3) ur_stencil(float* out, int steps, float seed) {
  constexpr int RB = Q + 2, BW = 34;
  __shared__ __align__(16) float2 ring[RB * BW + 8];
  __shared__ unsigned stage[256 * 8];
  for (int i = threadIdx.x; i < RB * BW + 8; i += blockDim.x) ring[i] = make_float2(seed * i, seed);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pl = MODE == 3 ? (threadIdx.x >> 5) : __shfl_sync(0xffffffffu, warp, 0);
  float2 wv[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) wv[i] = c_w[pl * 28 + i];
  if (MODE == 3) {
#pragma unroll
    for (int i = 0; i < 27; ++i) { wv[i].x += lane * 1e-9f; }      // make them per-lane (vector registers)
  }
  const float2 bia = c_w[pl * 28 + 27];
  float2 acc[3][Q];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[a][q] = make_float2(0.f, 0.f);
  const unsigned base = (unsigned)__cvta_generic_to_shared(ring) + lane * 8;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(stage) + threadIdx.x * 4;
  float2 xr = make_float2(seed, seed * 2.f);
  float2 ssum = make_float2(0.f, 0.f);
  for (int s3 = 0; s3 < steps; s3 += 3) {
#pragma unroll
   for (int u = 0; u < 3; ++u) {
    const int s = s3 + u;
    if (MODE == 4) {
      unsigned long long ca;
      asm("cvta.to.const.u64 %0, %1;" : "=l"(ca) : "l"((unsigned long long)(c_w + pl * 28)));
#pragma unroll
      for (int i = 0; i < 27; ++i)
        asm volatile("ld.const.v2.f32 {%0, %1}, [%2];" : "=f"(wv[i].x), "=f"(wv[i].y) : "l"(ca + i * 8u));
    }
#pragma unroll
    for (int rr = 0; rr < RB; ++rr) {
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) {
        float2 x;
        if (MODE == 0) { x = xr; xr.x += 1.0f; }
        else x = lds2(base + (rr * BW + dw) * 8);
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          const int q = rr - dh;
          if (q >= 0 && q < Q) {
            acc[0][q] = __ffma2_rn(x, wv[(0 * 3 + dh) * 3 + dw], acc[0][q]);
            acc[1][q] = __ffma2_rn(x, wv[(1 * 3 + dh) * 3 + dw], acc[1][q]);
            acc[2][q] = __ffma2_rn(x, wv[(2 * 3 + dh) * 3 + dw], acc[2][q]);
          }
        }
      }
    }
    if (MODE == 2 || MODE == 5 || MODE == 6) {
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        float2 v = __fadd2_rn(MODE == 2 ? acc[s % 3 == 0 ? 0 : 1][q] : acc[u][q], bia);
        if (MODE != 5) {
          const float2 h = __fmul2_rn(v, make_float2(0.5f, 0.5f));
          float2 t;
          asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
          v = __ffma2_rn(h, t, h);
        }
        const __nv_bfloat162 hb = __float22bfloat162_rn(v);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + q * 1024), "r"(*reinterpret_cast<const unsigned*>(&hb)) : "memory");
        ssum = __fadd2_rn(ssum, v);
      }
    }
   }
  }
  float r = ssum.x + ssum.y + xr.x;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int q = 0; q < Q; ++q) r += acc[a][q].x + acc[a][q].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  float2 hw[8 * 28];
  for (int i = 0; i < 8 * 28; ++i) hw[i] = make_float2(1e-3f * i, -1e-3f * i);
  cudaMemcpyToSymbol(c_w, hw, sizeof(hw));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int steps = 1998;
  const char* nm[7] = {"FFMA2 R,R,UR,R  x in registers", "FFMA2 R,R,UR,R  + 30 LDS.64 / 216", "  + epilogue (swish, pack, STS) per step, dynamic set",
                       "FFMA2 R,R,R,R   + 30 LDS.64 / 216 (taps in vector registers)", "  (1) + 27 LDCU.64 per step",
                       "  (1) + epilogue without swish, static set", "  (1) + epilogue with swish, static set"};
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int mode = 0; mode < 7; ++mode)
      for (int rep = 0; rep < 2; ++rep) {
        const int bl = p.multiProcessorCount * ctas, th = 256;
        cudaEventRecord(e0);
        if (mode == 0) ur_stencil<0, 8><<<bl, th>>>(out, steps, 1e-3f);
        if (mode == 1) ur_stencil<1, 8><<<bl, th>>>(out, steps, 1e-3f);
        if (mode == 2) ur_stencil<2, 8><<<bl, th>>>(out, steps, 1e-3f);
        if (mode == 3) ur_stencil<3, 8><<<bl, th>>>(out, steps, 1e-3f);
        if (mode == 4) ur_stencil<4, 8><<<bl, th>>>(out, steps, 1e-3f);
        if (mode == 5) ur_stencil<5, 8><<<bl, th>>>(out, steps, 1e-3f);
        if (mode == 6) ur_stencil<6, 8><<<bl, th>>>(out, steps, 1e-3f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double cyc = ms * 1e-3 * p.clockRate * 1e3;
        const double per_smsp = (double)bl * th / 32 * steps * 216 / p.multiProcessorCount / 4;
        if (rep) printf("%2d warps/SM  %-62s %7.3f ms  %.2f cycles per FFMA2 per scheduler (nominal clock)\n", ctas * 8, nm[mode], ms, cyc / per_smsp);
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
