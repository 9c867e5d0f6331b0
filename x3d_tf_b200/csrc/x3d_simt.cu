// CUDA-core (SIMT) kernels of the X3D forward path for sm_100a:
//   stem (fused conv_s + conv_t + BN + ReLU), SE MLP, global average pool, softmax + view mean, and the generic pointwise
//   GEMM behind x3d_pw_fwd: fp32 activations (fp32 path, training step) on the tensor cores with the 3xTF32 split
//   (pw_gemm_tf32x3_kernel), bf16 activations without the tcgen05 path's layout requirements on CUDA cores (pw_gemm_kernel).
// The bf16 tensor-core pointwise GEMM lives in x3d_pw_tc.cu, the channelwise stencil in x3d_dw_tma.cu.
#include "common.cuh"

namespace x3d {

// =====================================================================================
// Stem.  One CTA = one 8x16 output tile of one clip (and one chunk of <=32 channels), marching
// over the T frames: per frame the 17x33x3 input patch is staged in shared memory, every thread
// evaluates conv_s for its channel pair on up to 8 pixels (weights live in registers), pushes the
// result into a 5-deep register ring and emits conv_t + BN + ReLU for frame t-2.  Input is read
// once, output written once; conv_s results never leave registers.
constexpr int kStemTH = 8, kStemTW = 16, kStemPix = kStemTH * kStemTW;
constexpr int kStemIH = 2 * kStemTH + 1, kStemIW3 = (2 * kStemTW + 1) * 3, kStemPitch = 100;
constexpr int kStemThreads = 256, kStemNP = 8, kStemKT = 5;

template <typename TI, typename TO>
__global__ void __launch_bounds__(kStemThreads)
stem_kernel(const TI* __restrict__ in, const float* __restrict__ ws, const float* __restrict__ wt,
            const float* __restrict__ bias, TO* __restrict__ out, int T, int H, int W, int Ho,
            int Wo, int C, int cchunks) {
  __shared__ float s_in[kStemIH * kStemPitch];
  const int tid = threadIdx.x;
  const int n = blockIdx.z / cchunks, cc = blockIdx.z % cchunks;
  const int c_lo = cc * 32;
  const int c2c = (min(C, c_lo + 32) - c_lo) >> 1;          // channel pairs in this chunk
  const int PL = kStemThreads / c2c;                         // pixel lanes
  const int cp = tid % c2c, pl = tid / c2c;
  const bool lane_on = pl < PL;
  const int c = c_lo + 2 * cp;
  const int ho0 = blockIdx.y * kStemTH, wo0 = blockIdx.x * kStemTW;

  float2 wsr[27], wtr[kStemKT];
#pragma unroll
  for (int i = 0; i < 27; ++i) wsr[i] = ld2(ws + i * C + c);
#pragma unroll
  for (int i = 0; i < kStemKT; ++i) wtr[i] = ld2(wt + i * C + c);
  const float2 b = ld2(bias + c);

  int soff[kStemNP];          // smem offset of each pixel's patch; -1 = pixel not in tile
  long ooff[kStemNP];         // output offset inside one frame; -1 = outside the image
#pragma unroll
  for (int i = 0; i < kStemNP; ++i) {
    const int pix = pl + i * PL;
    const int py = pix / kStemTW, px = pix % kStemTW;
    const bool in_tile = lane_on && pix < kStemPix;
    soff[i] = in_tile ? (2 * py) * kStemPitch + 2 * px * 3 : -1;
    const int ho = ho0 + py, wo = wo0 + px;
    ooff[i] = (in_tile && ho < Ho && wo < Wo) ? ((long)ho * Wo + wo) * C + c : -1;
  }

  float2 ring[kStemKT][kStemNP];
#pragma unroll
  for (int d = 0; d < kStemKT; ++d)
#pragma unroll
    for (int i = 0; i < kStemNP; ++i) ring[d][i] = make_float2(0.f, 0.f);

  const long in_frame = (long)H * W * 3;
  const TI* in_n = in + (long)n * T * in_frame;
  TO* out_n = out + (long)n * T * Ho * Wo * C;
  const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 1;

  for (int t = 0; t < T + kStemKT / 2; ++t) {
    if (t < T) {
      const TI* fr = in_n + (long)t * in_frame;
      for (int idx = tid; idx < kStemIH * kStemIW3; idx += kStemThreads) {
        const int r = idx / kStemIW3, e = idx - r * kStemIW3;
        const int hi = hi0 + r, wi = wi0 + e / 3;
        float v = 0.f;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) v = ld1(fr + ((long)hi * W + wi0) * 3 + e);
        s_in[r * kStemPitch + e] = v;
      }
    }
    __syncthreads();
    // shift the ring, then push conv_s(frame t) (zero beyond the clip: temporal zero padding)
#pragma unroll
    for (int d = 0; d + 1 < kStemKT; ++d)
#pragma unroll
      for (int i = 0; i < kStemNP; ++i) ring[d][i] = ring[d + 1][i];
#pragma unroll
    for (int i = 0; i < kStemNP; ++i) {
      float2 acc = make_float2(0.f, 0.f);
      if (t < T && soff[i] >= 0) {
        const float* p = s_in + soff[i];
#pragma unroll
        for (int dh = 0; dh < 3; ++dh)
#pragma unroll
          for (int q = 0; q < 9; ++q) {
            const float x = p[dh * kStemPitch + q];
            acc = fma2(make_float2(x, x), wsr[dh * 9 + q], acc);
          }
      }
      ring[kStemKT - 1][i] = acc;
    }
    const int to = t - kStemKT / 2;
    if (to >= 0) {
      TO* of = out_n + (long)to * Ho * Wo * C;
#pragma unroll
      for (int i = 0; i < kStemNP; ++i) {
        if (ooff[i] < 0) continue;
        float2 y = b;
#pragma unroll
        for (int d = 0; d < kStemKT; ++d) y = fma2(ring[d][i], wtr[d], y);
        y.x = fmaxf(y.x, 0.f);
        y.y = fmaxf(y.y, 0.f);
        st2(of + ooff[i], y);
      }
    }
    __syncthreads();
  }
}

// =====================================================================================
// SE MLP: one CTA per clip.
__global__ void __launch_bounds__(256)
se_mlp_kernel(const float* __restrict__ partial, int nblk, float inv_count,
              const float* __restrict__ w1, const float* __restrict__ b1,
              const float* __restrict__ w2, const float* __restrict__ b2,
              float* __restrict__ scale, int C, int Cw) {
  extern __shared__ float sm[];      // mean[C] | z[Cw] | red[256]
  float* mean = sm;
  float* z = sm + C;
  float* red = z + Cw;
  const int n = blockIdx.x, tid = threadIdx.x;
  pdl_wait();                                   // (launch_pdl) the tile sums are the previous kernel's output
  pdl_trigger();
  const float* p = partial + (long)n * nblk * C;
  // tile sums -> mean: the tiles of a channel are split over KG thread groups (independent,
  // coalesced loads; up to 64 tiles at the stage-2 resolution), combined in fixed order
  const int CP = C <= 64 ? 64 : (C <= 128 ? 128 : 256), KG = 256 / CP;
  for (int cb = 0; cb < C; cb += CP) {
    const int c = cb + (tid % CP), kg = tid / CP;
    float a = 0.f;
    if (c < C) {
#pragma unroll 4
      for (int k = kg; k < nblk; k += KG) a += __ldg(p + (long)k * C + c);
    }
    red[tid] = a;
    __syncthreads();
    if (kg == 0 && c < C) {
      for (int g = 1; g < KG; ++g) a += red[g * CP + (tid % CP)];
      mean[c] = a * inv_count;
    }
    __syncthreads();
  }
  // fc1: thread tid walks w1 (row-major [C][Cw]) at tid, tid + 256, ...: coalesced, independent loads
  // (a warp per output column with a strided walk over C was a chain of ~14 exposed L2 round trips per
  // column: 20 of the kernel's 21 us).  Column j = tid % Cw is fixed per thread when Cw divides 256;
  // the 256 / Cw partial sums of a column meet in shared memory.
  if (256 % Cw == 0) {
    const int j = tid % Cw, G = 256 / Cw;
    const int total = C * Cw;
    float a = 0.f;
#pragma unroll 8
    for (int e = tid; e < total; e += 256) a = fmaf(mean[e / Cw], __ldg(w1 + e), a);
    red[tid] = a;
    __syncthreads();
    if (tid < Cw) {
      float zsum = 0.f;
      for (int g = 0; g < G; ++g) zsum += red[g * Cw + j];
      z[j] = fmaxf(zsum + b1[j], 0.f);
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    for (int j = warp; j < Cw; j += nwarp) {
      float a = 0.f;
      for (int c = lane; c < C; c += 32) a = fmaf(mean[c], w1[(long)c * Cw + j], a);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0) z[j] = fmaxf(a + b1[j], 0.f);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) {
    float a = b2[c];
#pragma unroll 8
    for (int j = 0; j < Cw; ++j) a = fmaf(z[j], __ldg(w2 + (long)j * C + c), a);
    scale[(long)n * C + c] = 1.f / (1.f + expf(-a));
  }
}

// =====================================================================================
// Global average pool [N,P,C] -> [N,C] fp32.  grid (C/64 chunks, N); 256 threads = 4 row groups
// x 64 channels; fixed-order reduction.
template <typename T>
__global__ void __launch_bounds__(256)
avgpool_kernel(const T* __restrict__ in, float* __restrict__ out, long P, int C) {
  __shared__ float red[4][64];
  const int n = blockIdx.y, c = blockIdx.x * 64 + (threadIdx.x & 63), g = threadIdx.x >> 6;
  float a = 0.f;
  if (c < C) {
    const T* p = in + (long)n * P * C + c;
    for (long r = g; r < P; r += 4) a += ld1(p + r * C);
  }
  red[g][threadIdx.x & 63] = a;
  __syncthreads();
  if (g == 0 && c < C) {
    const int l = threadIdx.x;
    out[(long)n * C + c] = (red[0][l] + red[1][l] + red[2][l] + red[3][l]) / (float)P;
  }
}

// =====================================================================================
// Softmax over classes (fp32) and mean over the num_preds consecutive clips of a video.
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, u) : v + u;
  }
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}

__global__ void __launch_bounds__(256)
softmax_viewmean_kernel(const float* __restrict__ logits, float* __restrict__ probs, int ncls,
                        int num_preds) {
  __shared__ float sh[8];
  const int vid = blockIdx.x;
  const float inv = 1.f / (float)num_preds;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) probs[(long)vid * ncls + c] = 0.f;
  for (int v = 0; v < num_preds; ++v) {
    const float* row = logits + ((long)vid * num_preds + v) * ncls;
    float m = -INFINITY;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) m = fmaxf(m, row[c]);
    m = block_reduce(m, sh, true);
    float s = 0.f;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) s += expf(row[c] - m);
    s = block_reduce(s, sh, false);
    const float k = inv / s;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x)
      probs[(long)vid * ncls + c] += expf(row[c] - m) * k;
  }
}

// =====================================================================================
// Generic pointwise GEMM, fp32 accumulate on CUDA cores.  128x64 output tile, BK = 16,
// 256 threads, 8x4 outputs per thread.
constexpr int kBM = 128, kBN = 64, kBK = 16, kAPitch = kBM + 4;

struct PwParams {
  const void* A; const float* Wt; const float* bias; const void* R; const float* se; void* D;
  long M; int K, Nc, lda, ldw, ldr, ldd;
  long rows_per_clip;
  int swish, relu, gather, T, Ho, Wo, Hi, Wi, stride;
};

template <typename TA, typename TD, bool kFast>
__global__ void __launch_bounds__(256)
pw_gemm_kernel(const PwParams p) {
  __shared__ __align__(16) float As[kBK][kAPitch];
  __shared__ __align__(16) float Ws[kBK][kBN];
  const int tid = threadIdx.x;
  const long m0 = (long)blockIdx.x * kBM;
  const int n0 = blockIdx.y * kBN;
  const TA* A = static_cast<const TA*>(p.A);

  // loader role: row lr, k-half lk (8 consecutive k)
  const int lr = tid >> 1, lk = (tid & 1) * 8;
  const long lm = m0 + lr;
  long arow = -1;
  long clip = 0;
  if (lm < p.M) {
    arow = lm;
    if (p.gather) {
      long q = lm;
      const int wo = (int)(q % p.Wo); q /= p.Wo;
      const int ho = (int)(q % p.Ho); q /= p.Ho;       // q = n*T + t
      arow = (q * p.Hi + (long)ho * p.stride) * p.Wi + (long)wo * p.stride;
    }
    if (p.se) clip = lm / p.rows_per_clip;
  }
  // W loader role
  const int wk = tid >> 4, wn = (tid & 15) * 4;

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += kBK) {
    float av[8];
    const int ka = k0 + lk;
    if (arow >= 0 && ka < p.K) {
      ld8(A + arow * p.lda + ka, av);
      if (p.se) {
        const float* sp = p.se + clip * p.K + ka;
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] *= __ldg(sp + i);
      }
      if (p.swish) {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = av[i] * sigmoidf_<kFast>(av[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = 0.f;
    }
    float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + wk < p.K && n0 + wn < p.Nc)
      wv = __ldg(reinterpret_cast<const float4*>(p.Wt + (long)(k0 + wk) * p.ldw + n0 + wn));
    __syncthreads();          // previous tile fully consumed
#pragma unroll
    for (int i = 0; i < 8; ++i) As[lk + i][lr] = av[i];
    *reinterpret_cast<float4*>(&Ws[wk][wn]) = wv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }

  const int col = n0 + tx * 4;
  if (col >= p.Nc) return;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col));
    bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
  }
  TD* D = static_cast<TD*>(p.D);
  const TD* R = static_cast<const TD*>(p.R);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    float y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = acc[i][j] + bv[j];
    if (R) {
      float rv[4];
      ld4(R + m * p.ldr + col, rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] += rv[j];
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = fmaxf(y[j], 0.f);
    }
    st4(D + m * p.ldd + col, y);
  }
}

// =====================================================================================
// fp32 pointwise GEMM on the tensor cores: 3xTF32 split (x = hi + lo, both TF32; a.b ~ lo.hi + hi.lo
// + hi.hi accumulated in fp32), which keeps fp32-level accuracy (the 1e-4 logit bound of the fp32
// path and the gradient checks of the training step) at a third of the TF32 rate - still far above
// the FFMA rate the CUDA-core kernel above is bound by.  Same contract as pw_gemm_kernel<float,
// float>: 128x64 tile, BK = 16, 256 threads; 8 warps own 32x32 sub-tiles (2x4 m16n8k8 fragments).
// Two shared-memory stages, the next K slice is fetched into registers while the MMAs run.
constexpr int kTAPitch = kBK + 4;      // A tile [m][k]: fragment reads (8 rows x 4 k) hit 32 banks
constexpr int kTWPitch = kBN + 8;      // W tile [k][n]: fragment reads (4 k x 8 n) hit 32 banks

// hi = x rounded to TF32 (half away from zero) with integer ops, lo = x - hi (exact) pre-biased by half a
// TF32 ulp so that the tensor core's truncation of the low 13 mantissa bits rounds it.  cvt.rna.tf32
// runs on the 16-lane conversion pipe and made the kernel conversion-bound (measured: 124 MAC/clk/SM).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// TN = 64: 4 x 2 warps of 32 x 32; TN = 32 (Nc <= 32, e.g. the 24-channel projections): 8 x 1 warps of
// 16 x 32, so that no warp sits on columns beyond Nc.
template <int TN>
__global__ void __launch_bounds__(256, 2)
pw_gemm_tf32x3_kernel(const PwParams p) {
  constexpr int MF = TN == 64 ? 2 : 1, WPitch = TN + 8;
  __shared__ __align__(16) float As[2][kBM][kTAPitch];
  __shared__ __align__(16) float Ws[2][kBK][WPitch];
  const int tid = threadIdx.x;
  const long m0 = (long)blockIdx.x * kBM;
  const int n0 = blockIdx.y * TN;
  const float* A = static_cast<const float*>(p.A);

  // loader roles as in pw_gemm_kernel: A row lr / k-half lk, W row wk / 4 columns from wn
  const int lr = tid >> 1, lk = (tid & 1) * 8;
  const long lm = m0 + lr;
  long arow = -1;
  long clip = 0;
  if (lm < p.M) {
    arow = lm;
    if (p.gather) {
      long q = lm;
      const int wo = (int)(q % p.Wo); q /= p.Wo;
      const int ho = (int)(q % p.Ho); q /= p.Ho;       // q = n*T + t
      arow = (q * p.Hi + (long)ho * p.stride) * p.Wi + (long)wo * p.stride;
    }
    if (p.se) clip = lm / p.rows_per_clip;
  }
  const int wk = tid >> 4, wn = (tid & 15) * 4;

  float av[8];
  float4 wv;
  auto fetch = [&](int k0) {
    const int ka = k0 + lk;
    if (arow >= 0 && ka < p.K) {
      ld8(A + arow * p.lda + ka, av);
      if (p.se) {
        const float* sp = p.se + clip * p.K + ka;
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] *= __ldg(sp + i);
      }
      if (p.swish) {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = av[i] * sigmoidf_<false>(av[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) av[i] = 0.f;
    }
    wv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k0 + wk < p.K && wn < TN && n0 + wn < p.Nc)
      wv = __ldg(reinterpret_cast<const float4*>(p.Wt + (long)(k0 + wk) * p.ldw + n0 + wn));
  };
  auto stage = [&](int buf) {
    *reinterpret_cast<float4*>(&As[buf][lr][lk]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&As[buf][lr][lk + 4]) = make_float4(av[4], av[5], av[6], av[7]);
    if (wn < TN) *reinterpret_cast<float4*>(&Ws[buf][wk][wn]) = wv;
  };

  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wm = TN == 64 ? (warp & 3) * 32 : warp * 16, wc = TN == 64 ? (warp >> 2) * 32 : 0;
  float acc[MF][4][4];
#pragma unroll
  for (int i = 0; i < MF; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  fetch(0);
  stage(0);
  __syncthreads();
  const int nk = (p.K + kBK - 1) / kBK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * kBK);          // in flight while the MMAs below run
#pragma unroll
    for (int ks = 0; ks < kBK; ks += 8) {
      if (kt * kBK + ks >= p.K) break;               // K is a multiple of 8: skip an all-zero half
      uint32_t ah[MF][4], al[MF][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int i = 0; i < MF; ++i) {
        const float* ap = &As[buf][wm + i * 16 + g][ks + t];
        split_tf32(ap[0], ah[i][0], al[i][0]);
        split_tf32(ap[8 * kTAPitch], ah[i][1], al[i][1]);
        split_tf32(ap[4], ah[i][2], al[i][2]);
        split_tf32(ap[8 * kTAPitch + 4], ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* bp = &Ws[buf][ks + t][wc + j * 8 + g];
        split_tf32(bp[0], bh[j][0], bl[j][0]);
        split_tf32(bp[4 * WPitch], bh[j][1], bl[j][1]);
      }
      // term-major order: 8 independent MMAs between two that share an accumulator; column
      // fragments beyond Nc (warp-uniform) are skipped
#pragma unroll
      for (int term = 0; term < 3; ++term)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n0 + wc + j * 8 >= p.Nc) continue;
#pragma unroll
          for (int i = 0; i < MF; ++i) mma_tf32(acc[i][j], term == 0 ? al[i] : ah[i], term == 1 ? bl[j] : bh[j]);
        }
    }
    if (kt + 1 < nk) stage(buf ^ 1);                 // last read before the barrier that ended kt-1
    __syncthreads();
  }

  float* D = static_cast<float*>(p.D);
  const float* R = static_cast<const float*>(p.R);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = n0 + wc + j * 8 + 2 * t;         // Nc is a multiple of 4, col is even
    if (col >= p.Nc) continue;
    float2 bv = make_float2(0.f, 0.f);
    if (p.bias) bv = __ldg(reinterpret_cast<const float2*>(p.bias + col));
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long m = m0 + wm + i * 16 + g + h * 8;
        if (m >= p.M) continue;
        float2 y = make_float2(acc[i][j][2 * h] + bv.x, acc[i][j][2 * h + 1] + bv.y);
        if (R) {
          const float2 rv = __ldg(reinterpret_cast<const float2*>(R + m * p.ldr + col));
          y.x += rv.x; y.y += rv.y;
        }
        if (p.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); }
        *reinterpret_cast<float2*>(D + m * p.ldd + col) = y;
      }
  }
}

// =====================================================================================
// Row gather for the strided shortcut conv (ResBlock.residual: 1x1x1, stride (1,s,s), 'valid',
// model.py:360-367): copies the sampled pixels (n, t, ho*s, wo*s) into a dense [M_out, C] matrix
// so that the tensor-core GEMM can consume them through a plain 2-D TMA map.  16-byte vectors,
// grid-stride; reads 1/s^2 of the input rows, each exactly once.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long rows_out, int Ho,
                   int Wo, int Hi, int Wi, int stride, int vec_per_row) {
  const long total = rows_out * vec_per_row;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    long m = i / vec_per_row;
    const int v = (int)(i - m * vec_per_row);
    const int wo = (int)(m % Wo); m /= Wo;
    const int ho = (int)(m % Ho); m /= Ho;            // m = n*T + t
    const long src = (m * Hi + (long)ho * stride) * Wi + (long)wo * stride;
    out[i] = __ldg(in + src * vec_per_row + v);
  }
}

// =====================================================================================
// Skinny GEMM for the head (fc1 + ReLU, fc2 + bias; model.py:119-121): M = clips (tens), so the
// work is streaming the fp32 weights once.  CTA = 16 output columns x up to 128 rows; K is walked
// in chunks of 64 staged in shared memory; thread = (column, row group of 8).
constexpr int kSkK = 64, kSkM = 128;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cluster_rank(uint32_t smem_addr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// Small-M GEMM of the head (M = clips in flight, <= a few hundred rows): latency-bound, so the K range
// is split over the CTAs of a thread-block cluster (cluster dim z = splits), every CTA walks only
// K/splits (4 chunks of 64 at fc2's K = 2048 with 8 splits instead of 32 one after the other) and
// reads only its K slice of A; the partial tiles meet in the first CTA's epilogue through distributed
// shared memory, summed in rank order (deterministic).  SKN output columns per CTA (16, or 4 when 16
// would leave most SMs idle).
template <int SKN>
__global__ void __launch_bounds__(256)
skinny_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Wt,
                   const float* __restrict__ bias, float* __restrict__ D, int M, int K, int Nc,
                   int lda, int ldw, int ldd, int relu, int k_per) {
  constexpr int kGroups = 256 / SKN, kRows = kSkM / kGroups;     // row groups, rows per thread
  __shared__ __align__(16) float As[kSkM][kSkK + 4];
  __shared__ __align__(16) float Ws[kSkK][SKN];
  __shared__ __align__(16) float Part[kSkM][SKN];                // this CTA's partial tile (read by rank 0)
  const int tid = threadIdx.x, c = tid % SKN, g = tid / SKN;
  const int n0 = blockIdx.x * SKN, m0 = blockIdx.y * kSkM;
  const int rows = min(kSkM, M - m0);
  const int splits = gridDim.z, kz = blockIdx.z;                 // cluster = the splits of one tile
  const int k_lo = kz * k_per, k_hi = min(K, k_lo + k_per);
  float acc[kRows];
#pragma unroll
  for (int i = 0; i < kRows; ++i) acc[i] = 0.f;
  for (int k0 = k_lo; k0 < k_hi; k0 += kSkK) {
    __syncthreads();
    for (int i = tid; i < kSkM * (kSkK / 4); i += 256) {
      const int r = i / (kSkK / 4), kv = (i - r * (kSkK / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows && k0 + kv < k_hi) v = __ldg(reinterpret_cast<const float4*>(A + (long)(m0 + r) * lda + k0 + kv));
      *reinterpret_cast<float4*>(&As[r][kv]) = v;
    }
    for (int i = tid; i < kSkK * (SKN / 4); i += 256) {
      const int k = i / (SKN / 4), nv = (i - k * (SKN / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < k_hi && n0 + nv < Nc) v = __ldg(reinterpret_cast<const float4*>(Wt + (long)(k0 + k) * ldw + n0 + nv));
      *reinterpret_cast<float4*>(&Ws[k][nv]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < kSkK; k += 4) {
      const float w0 = Ws[k][c], w1 = Ws[k + 1][c], w2 = Ws[k + 2][c], w3 = Ws[k + 3][c];
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(&As[g + kGroups * i][k]);
        acc[i] = fmaf(a.x, w0, acc[i]);
        acc[i] = fmaf(a.y, w1, acc[i]);
        acc[i] = fmaf(a.z, w2, acc[i]);
        acc[i] = fmaf(a.w, w3, acc[i]);
      }
    }
  }
  if (splits > 1) {
    if (kz != 0) {
#pragma unroll
      for (int i = 0; i < kRows; ++i) Part[g + kGroups * i][c] = acc[i];
    }
    cluster_sync();                                              // partial tiles written
    if (kz == 0) {
      const uint32_t mine = smem_u32(&Part[0][0]);
      for (int r = 1; r < splits; ++r) {
        const uint32_t remote = map_to_cluster_rank(mine, r);
#pragma unroll
        for (int i = 0; i < kRows; ++i)
          acc[i] += ld_cluster_f32(remote + ((g + kGroups * i) * SKN + c) * 4);
      }
    }
    cluster_sync();                                              // nobody leaves while its tile is being read
    if (kz != 0) return;
  }
  const int col = n0 + c;
  if (col >= Nc) return;
  const float b = bias ? __ldg(bias + col) : 0.f;
#pragma unroll
  for (int i = 0; i < kRows; ++i) {
    const int r = g + kGroups * i;
    if (r < rows) {
      float y = acc[i] + b;
      if (relu) y = fmaxf(y, 0.f);
      D[(long)(m0 + r) * ldd + col] = y;
    }
  }
}

// --------------------------------------------------------------------------- host side
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

}  // namespace x3d

using namespace x3d;

extern "C" {

int x3d_stem_fwd(const void* in, int in_dtype, const float* ws, const float* wt,
                 const float* bias, void* out, int out_dtype, int N, int T, int H, int W, int C,
                 int kt, void* stream) {
  X3D_REQUIRE(in && ws && wt && bias && out, X3D_ERR_INVALID_ARG, "x3d_stem_fwd: null pointer");
  X3D_REQUIRE(kt == kStemKT, X3D_ERR_UNSUPPORTED, "x3d_stem_fwd: temporal filter %d (only 5)", kt);
  X3D_REQUIRE(C > 0 && C % 8 == 0, X3D_ERR_INVALID_ARG, "x3d_stem_fwd: C=%d not a multiple of 8", C);
  X3D_REQUIRE(N > 0 && T > 0 && H > 0 && W > 0, X3D_ERR_INVALID_ARG, "x3d_stem_fwd: empty input");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int cchunks = (C + 31) / 32;
  X3D_REQUIRE((long)N * cchunks <= 65535, X3D_ERR_UNSUPPORTED, "x3d_stem_fwd: batch too large for one launch");
  dim3 grid((Wo + kStemTW - 1) / kStemTW, (Ho + kStemTH - 1) / kStemTH, N * cchunks);
  cudaStream_t st = S(stream);
#define X3D_STEM(TI, TO)                                                                         \
  stem_kernel<TI, TO><<<grid, kStemThreads, 0, st>>>(static_cast<const TI*>(in), ws, wt, bias,   \
                                                     static_cast<TO*>(out), T, H, W, Ho, Wo, C,  \
                                                     cchunks)
  if (in_dtype == X3D_F32 && out_dtype == X3D_F32) X3D_STEM(float, float);
  else if (in_dtype == X3D_F32 && out_dtype == X3D_BF16) X3D_STEM(float, bf16);
  else if (in_dtype == X3D_BF16 && out_dtype == X3D_BF16) X3D_STEM(bf16, bf16);
  else if (in_dtype == X3D_BF16 && out_dtype == X3D_F32) X3D_STEM(bf16, float);
  else X3D_REQUIRE(false, X3D_ERR_INVALID_ARG, "x3d_stem_fwd: bad dtype %d/%d", in_dtype, out_dtype);
#undef X3D_STEM
  return check_launch("x3d_stem_fwd");
}

int x3d_se_mlp_fwd(const float* partial, int nblk, float inv_count, const float* w1,
                   const float* b1, const float* w2, const float* b2, float* scale, int N, int C,
                   int Cw, void* stream) {
  X3D_REQUIRE(partial && w1 && b1 && w2 && b2 && scale, X3D_ERR_INVALID_ARG, "x3d_se_mlp_fwd: null pointer");
  X3D_REQUIRE(N > 0 && C > 0 && Cw > 0 && Cw <= 64 && nblk > 0, X3D_ERR_INVALID_ARG, "x3d_se_mlp_fwd: bad size");
  const cudaError_t le = launch_pdl(se_mlp_kernel, dim3(N), dim3(256), sizeof(float) * (C + Cw + 256), S(stream), partial,
                                    nblk, inv_count, w1, b1, w2, b2, scale, C, Cw);
  X3D_REQUIRE(le == cudaSuccess, X3D_ERR_LAUNCH, "x3d_se_mlp_fwd: launch: %s", cudaGetErrorString(le));
  return check_launch("x3d_se_mlp_fwd");
}

int x3d_avgpool_fwd(const void* in, float* out, int N, int64_t P, int C, int dtype, void* stream) {
  X3D_REQUIRE(in && out && N > 0 && N <= 65535 && P > 0 && C > 0, X3D_ERR_INVALID_ARG, "x3d_avgpool_fwd: bad argument");
  dim3 grid((C + 63) / 64, N);
  if (dtype == X3D_BF16)
    avgpool_kernel<bf16><<<grid, 256, 0, S(stream)>>>(static_cast<const bf16*>(in), out, P, C);
  else if (dtype == X3D_F32)
    avgpool_kernel<float><<<grid, 256, 0, S(stream)>>>(static_cast<const float*>(in), out, P, C);
  else
    X3D_REQUIRE(false, X3D_ERR_INVALID_ARG, "x3d_avgpool_fwd: dtype %d", dtype);
  return check_launch("x3d_avgpool_fwd");
}

int x3d_softmax_viewmean_fwd(const float* logits, float* probs, int N, int ncls, int num_preds,
                             void* stream) {
  X3D_REQUIRE(logits && probs && N > 0 && ncls > 0 && num_preds > 0, X3D_ERR_INVALID_ARG, "x3d_softmax_viewmean_fwd: bad argument");
  X3D_REQUIRE(N % num_preds == 0, X3D_ERR_INVALID_ARG, "x3d_softmax_viewmean_fwd: batch %d is not a multiple of num_preds %d", N, num_preds);
  softmax_viewmean_kernel<<<N / num_preds, 256, 0, S(stream)>>>(logits, probs, ncls, num_preds);
  return check_launch("x3d_softmax_viewmean_fwd");
}

int x3d_pw_fwd(const x3d_pw_args* a, void* stream) {
  X3D_REQUIRE(a && a->A && a->Wt && a->D, X3D_ERR_INVALID_ARG, "x3d_pw_fwd: null pointer");
  X3D_REQUIRE(a->M > 0 && a->K > 0 && a->Nc > 0, X3D_ERR_INVALID_ARG, "x3d_pw_fwd: empty problem");
  X3D_REQUIRE(a->K % 8 == 0 && a->lda % 8 == 0, X3D_ERR_INVALID_ARG, "x3d_pw_fwd: K=%d/lda=%d must be multiples of 8", a->K, a->lda);
  X3D_REQUIRE(a->Nc % 4 == 0 && a->ldw % 4 == 0 && a->ldd % 4 == 0 && (!a->R || a->ldr % 4 == 0), X3D_ERR_INVALID_ARG, "x3d_pw_fwd: Nc/ldw/ldd/ldr must be multiples of 4");
  X3D_REQUIRE(!a->se || a->rows_per_clip > 0, X3D_ERR_INVALID_ARG, "x3d_pw_fwd: rows_per_clip missing");
  X3D_REQUIRE(!a->gather || (a->stride >= 1 && a->Ho > 0 && a->Wo > 0 && a->Hi > 0 && a->Wi > 0), X3D_ERR_INVALID_ARG, "x3d_pw_fwd: bad gather geometry");
  PwParams p;
  p.A = a->A; p.Wt = a->Wt; p.bias = a->bias; p.R = a->R; p.se = a->se; p.D = a->D;
  p.M = a->M; p.K = a->K; p.Nc = a->Nc; p.lda = a->lda; p.ldw = a->ldw; p.ldr = a->ldr; p.ldd = a->ldd;
  p.rows_per_clip = a->rows_per_clip; p.swish = a->swish; p.relu = a->relu; p.gather = a->gather;
  p.T = a->T; p.Ho = a->Ho; p.Wo = a->Wo; p.Hi = a->Hi; p.Wi = a->Wi; p.stride = a->stride;
  const long mt = (a->M + kBM - 1) / kBM;
  X3D_REQUIRE(mt <= 2147483647L, X3D_ERR_UNSUPPORTED, "x3d_pw_fwd: M too large");
  dim3 grid((unsigned)mt, (a->Nc + kBN - 1) / kBN);
  cudaStream_t st = S(stream);
  if (a->a_dtype == X3D_F32 && a->d_dtype == X3D_F32)
  {
    if (a->Nc <= 32) {
      grid.y = 1;
      pw_gemm_tf32x3_kernel<32><<<grid, 256, 0, st>>>(p);
    } else {
      pw_gemm_tf32x3_kernel<64><<<grid, 256, 0, st>>>(p);
    }
  }
  else if (a->a_dtype == X3D_BF16 && a->d_dtype == X3D_BF16)
    pw_gemm_kernel<bf16, bf16, true><<<grid, 256, 0, st>>>(p);
  else if (a->a_dtype == X3D_BF16 && a->d_dtype == X3D_F32)
    pw_gemm_kernel<bf16, float, true><<<grid, 256, 0, st>>>(p);
  else
    X3D_REQUIRE(false, X3D_ERR_INVALID_ARG, "x3d_pw_fwd: unsupported dtype pair %d/%d", a->a_dtype, a->d_dtype);
  return check_launch("x3d_pw_fwd");
}

int x3d_gather_rows_fwd(const void* in, void* out, int NT, int Hi, int Wi, int stride, int C,
                        int dtype, void* stream) {
  X3D_REQUIRE(in && out, X3D_ERR_INVALID_ARG, "x3d_gather_rows_fwd: null pointer");
  X3D_REQUIRE(NT > 0 && Hi > 0 && Wi > 0 && stride >= 1 && C > 0, X3D_ERR_INVALID_ARG, "x3d_gather_rows_fwd: bad extent");
  X3D_REQUIRE(dtype == X3D_F32 || dtype == X3D_BF16, X3D_ERR_INVALID_ARG, "x3d_gather_rows_fwd: dtype %d", dtype);
  const int es = dtype == X3D_BF16 ? 2 : 4;
  X3D_REQUIRE((C * es) % 16 == 0, X3D_ERR_INVALID_ARG, "x3d_gather_rows_fwd: row of %d bytes is not a multiple of 16", C * es);
  X3D_REQUIRE(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_gather_rows_fwd: pointers must be 16-byte aligned");
  const int Ho = (Hi - 1) / stride + 1, Wo = (Wi - 1) / stride + 1;
  const long rows = (long)NT * Ho * Wo;
  const int vpr = C * es / 16;
  long blocks = (rows * vpr + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  gather_rows_kernel<<<(unsigned)blocks, 256, 0, S(stream)>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out),
                                                             rows, Ho, Wo, Hi, Wi, stride, vpr);
  return check_launch("x3d_gather_rows_fwd");
}

int x3d_head_fc_fwd(const float* A, const float* Wt, const float* bias, float* D, int M, int K,
                    int Nc, int lda, int ldw, int ldd, int relu, void* stream) {
  X3D_REQUIRE(A && Wt && D, X3D_ERR_INVALID_ARG, "x3d_head_fc_fwd: null pointer");
  X3D_REQUIRE(M > 0 && K > 0 && Nc > 0, X3D_ERR_INVALID_ARG, "x3d_head_fc_fwd: empty problem");
  X3D_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && Nc % 4 == 0, X3D_ERR_INVALID_ARG, "x3d_head_fc_fwd: K/lda/ldw/Nc must be multiples of 4");
  X3D_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Wt)) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_head_fc_fwd: pointers must be 16-byte aligned");
  const int my = (M + kSkM - 1) / kSkM;
  X3D_REQUIRE(my <= 65535, X3D_ERR_UNSUPPORTED, "x3d_head_fc_fwd: M too large");
  // K split over a cluster: chunks of 64, at most 8 CTAs, at least 2 chunks each
  int splits = (K + 2 * kSkK - 1) / (2 * kSkK);
  if (splits > 8) splits = 8;
  if (splits < 1) splits = 1;
  const int k_per = ((K + splits - 1) / splits + kSkK - 1) / kSkK * kSkK;
  splits = (K + k_per - 1) / k_per;
  const bool narrow = (long)((Nc + 15) / 16) * my * splits < 96;   // too few 16-column CTAs to cover the SMs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(narrow ? (Nc + 3) / 4 : (Nc + 15) / 16, my, splits);
  cfg.blockDim = dim3(256);
  cfg.stream = S(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = splits;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const cudaError_t le = narrow
      ? cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<4>, A, Wt, bias, D, M, K, Nc, lda, ldw, ldd, relu, k_per)
      : cudaLaunchKernelEx(&cfg, skinny_gemm_kernel<16>, A, Wt, bias, D, M, K, Nc, lda, ldw, ldd, relu, k_per);
  X3D_REQUIRE(le == cudaSuccess, X3D_ERR_LAUNCH, "x3d_head_fc_fwd: launch (cluster of %d): %s", splits, cudaGetErrorString(le));
  return check_launch("x3d_head_fc_fwd");
}

}  // extern "C"
