// Stem on the tensor cores (bf16 activations): X3D_Stem.call, reference model.py:202-210.
//
// conv_s (1x3x3, stride 2, 3->C) is linear and nothing non-linear sits between it and the
// channelwise temporal conv_t (kt x1x1), so the pair equals ONE dense kt x3x3 convolution with
//     Wc[dt][k][c] = ws[k][c] * wt[dt][c] * bn_scale[c],   k = (dh, dw, ci) in 0..26
// (SURVEY.md Appendix A.2).  As an implicit GEMM per output tile of 128 pixels:
//     D[128, C] = sum_{dt} A_{t-2+dt}[128, 32] * Wc[dt][32, C]
// where A_f is the im2col block of input frame f (27 taps, zero-padded to K = 32).  A CTA marches
// over the T frames of one 8x16 output tile: each frame's im2col block is built ONCE (27 strided
// loads per pixel, written K-major into a 6-deep shared-memory ring) and consumed by the 5 output
// frames it contributes to; one elected thread issues the tcgen05.mma's (M=128, N=32, K=16) into a
// double-buffered TMEM accumulator; the 4 worker warps read it back (tcgen05.ld), add the BN
// shift, apply ReLU and store 16-byte bf16 vectors.  conv_s results never exist in memory.
#include "tma_common.cuh"

namespace x3d {
namespace stemtc {

using namespace ptx;

constexpr int kTH = 8, kTW = 16, kPix = kTH * kTW;        // 128 output pixels = UMMA M
constexpr int kKT = 5, kRing = 6;
constexpr int kN = 32;                                     // UMMA N (C <= 32, zero padded)
constexpr int kABytes = kPix * 32 * 2;                     // one im2col block [4 kchunks][128][16B]
constexpr int kWBytes = kN * 32 * 2;                       // one weight block [4 kchunks][32][16B]
constexpr int kThreads = 160;                              // 4 worker warps + 1 MMA warp

__device__ __forceinline__ void tcgen05_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major operand without swizzle: 8x16B core matrices; LBO = byte stride between core matrices
// adjacent in K, SBO = byte stride between core matrices adjacent in M/N.
__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;                     // descriptor version (sm_100)
  return d;                                                // layout_type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(kPix >> 4) << 24);
}

__device__ __forceinline__ uint32_t to_bf16_bits(float x) {
  __nv_bfloat16 h = __float2bfloat16_rn(x);
  return static_cast<uint32_t>(*reinterpret_cast<unsigned short*>(&h));
}
// `lut` is only used by the uint8 input stage: [3][256] bf16 bits of ((u/nv) - mean[c]) / std[c]
__device__ __forceinline__ uint32_t load_bf16_bits(const float* p, const unsigned short*, int) {
  return to_bf16_bits(__ldg(p));
}
__device__ __forceinline__ uint32_t load_bf16_bits(const bf16* p, const unsigned short*, int) {
  return static_cast<uint32_t>(__ldg(reinterpret_cast<const unsigned short*>(p)));
}
__device__ __forceinline__ uint32_t load_bf16_bits(const uint8_t* p, const unsigned short* lut, int ch) {
  return static_cast<uint32_t>(lut[ch * 256 + __ldg(p)]);
}

// utils.normalize (reference utils.py:42-72): x / norm_value, then (x - mean) / std, fp32.
struct InputNorm {
  float mean[3], std[3], norm_value;
};

// kStaged (bf16 input, even W): the 17 x 33-pixel input patch of a frame is copied into a 3-deep
// shared-memory ring with 4-byte cp.async (7 per thread and frame, zero fill = the conv's padding) two
// frames ahead, and every worker assembles its 27 taps from 15 shared-memory words; otherwise every
// worker gathers its 27 taps with 2-byte global loads.
constexpr int kSW = 50, kSRows = 2 * kTH + 1, kSWords = kSW * kSRows, kSRing = 3, kSBytes = 3456;

template <typename TI, bool kStaged>
// (capping the registers for a third CTA per SM serialises the loads again: 0.675 ms vs 0.605 ms)
__global__ void __launch_bounds__(kThreads)
stem_tc_kernel(const TI* __restrict__ in, const uint4* __restrict__ wc, const float* __restrict__ bias,
               bf16* __restrict__ out, int T, int H, int W, int Ho, int Wo, int C, const InputNorm nrm) {
  extern __shared__ __align__(128) uint8_t stem_smem_raw[];
  const uint32_t raw_s = smem_u32(stem_smem_raw);
  const uint32_t smem_s = (raw_s + 127u) & ~127u;
  uint8_t* smem = stem_smem_raw + (smem_s - raw_s);
  // layout: [A ring kRing x 8 KB][W kKT x 2 KB][bias 32 f32][barriers][tmem slot][u8 LUT 3x256 u16]
  const uint32_t a_s = smem_s;
  const uint32_t w_s = a_s + kRing * kABytes;
  float* s_bias = reinterpret_cast<float*>(smem + kRing * kABytes + kKT * kWBytes);
  uint64_t* built = reinterpret_cast<uint64_t*>(s_bias + 32);     // [kRing] im2col block ready
  uint64_t* t_full = built + kRing;                               // [2] accumulator ready
  uint64_t* t_empty = t_full + 2;                                 // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  unsigned short* s_lut = reinterpret_cast<unsigned short*>(tmem_slot + 4);
  const uint32_t stage_s = smem_u32(s_lut + 768);          // [kSRing][kSBytes] input patches (kStaged)

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n = blockIdx.z;
  const int ho0 = blockIdx.y * kTH, wo0 = blockIdx.x * kTW;

  // one-time setup: zero the ring (K padding 27..31 must read as 0), weights, bias, barriers, TMEM
  for (int i = tid; i < kRing * kABytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < kKT * kWBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(smem + kRing * kABytes)[i] = __ldg(wc + i);
  if (tid < 32) s_bias[tid] = tid < C ? bias[tid] : 0.f;
  if (sizeof(TI) == 1) {
    // input stage fused into the loader: the 256 possible pixel values of each channel, normalised
    // in the reference's fp32 operation order and rounded to bf16 once
    for (int i = tid; i < 768; i += kThreads) {
      const int ch = i >> 8;
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i & 255), nrm.norm_value), nrm.mean[ch]),
                                nrm.std[ch]);
      s_lut[i] = static_cast<unsigned short>(to_bf16_bits(v));
    }
  }
  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) mbar_init(&built[i], kPix);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], kPix);
    }
    fence_barrier_init();
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(
                     smem_u32(tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();                     // generic-proxy writes (zeros, weights) -> async proxy
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ workers: one pixel each
    const int r = tid;
    const int py = r >> 4, px = r & 15;
    const int ho = ho0 + py, wo = wo0 + px;
    const bool pix_ok = ho < Ho && wo < Wo;
    const int hi = 2 * ho - 1, wi = 2 * wo - 1;
    bool rv[3], cv[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      rv[d] = pix_ok && hi + d >= 0 && hi + d < H;
      cv[d] = wi + d >= 0 && wi + d < W;
    }
    const int row_elems = W * 3;
    const long frame_elems = static_cast<long>(H) * row_elems;
    const TI* src0 = in + static_cast<long>(n) * T * frame_elems + static_cast<long>(hi) * row_elems + wi * 3;
    const uint32_t a_row = static_cast<uint32_t>(r) * 16;
    bf16* dst0 = out + ((static_cast<long>(n) * T * Ho + ho) * Wo + wo) * C;
    const long out_frame = static_cast<long>(Ho) * Wo * C;
    const int q = warp;                                    // TMEM lane quarter of this warp
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    auto epilogue = [&](int t) {
      // ---- epilogue of output frame t
      const int as = t & 1;
      mbar_wait(&t_full[as], static_cast<uint32_t>((t >> 1) & 1));
      tcgen05_after_sync();
      bf16* dst = dst0 + t * out_frame;
      // all TMEM reads of the frame first (one wait), so the accumulator is released before the stores
      uint32_t v[kN / 8][8];
#pragma unroll
      for (int c8 = 0; c8 < kN; c8 += 8)
        if (c8 < C) tmem_ld8(t_lane + as * kN + c8, v[c8 / 8]);          // uniform
      tmem_ld_wait();
      tcgen05_before_sync();
      mbar_arrive(&t_empty[as]);
      if (pix_ok) {
#pragma unroll
        for (int c8 = 0; c8 < kN; c8 += 8) {
          if (c8 < C) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              y[j] = fmaxf(__uint_as_float(v[c8 / 8][j]) + s_bias[c8 + j], 0.f);
            __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(y[2], y[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(y[4], y[5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(y[6], y[7]);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&p0);
            o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2);
            o.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(dst + c8) = o;
          }
        }
      }
    };

    if constexpr (kStaged) {
      // ---- staged loader: word idx = tid + 128 j of the [17][50]-word patch; its source offset inside
      // a frame (bytes) is fixed for the whole march, -1 = outside the image (zero = conv padding)
      const long row_bytes = static_cast<long>(W) * 6;
      const char* in_n = reinterpret_cast<const char*>(in) + static_cast<long>(n) * T * frame_elems * 2;
      long soff[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int idx = tid + j * 128;
        const int rr = idx / kSW, ww = idx - rr * kSW;
        const int h = 2 * ho0 - 1 + rr;
        const long off = 12L * wo0 - 8 + 4L * ww;          // word 0 starts 2 bytes before pixel 2*wo0-1
        soff[j] = (idx < kSWords && h >= 0 && h < H && off >= 0 && off < row_bytes) ? h * row_bytes + off : -1;
      }
      auto issue_frame = [&](int f) {
        if (f < T) {
          const char* fr = in_n + static_cast<long>(f) * frame_elems * 2;
          const uint32_t dstb = stage_s + (f % kSRing) * kSBytes + tid * 4;
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            if (tid + j * 128 < kSWords) {
              const bool ok = soff[j] >= 0;
              const char* src = ok ? fr + soff[j] : reinterpret_cast<const char*>(in);
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dstb + j * 512), "l"(src),
                           "r"(ok ? 4 : 0)
                           : "memory");
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // one group per frame, empty past the end
      };
      issue_frame(0);
      issue_frame(1);
      for (int i = 0; i < T + 3; ++i) {
        if (i < T) {
          asm volatile("cp.async.wait_group 1;" ::: "memory");   // frame i has landed (i+1 may be in flight)
          asm volatile("bar.sync 1, 128;" ::: "memory");         // ... for every worker; frame i-1 fully read
          const uint32_t sb = stage_s + (i % kSRing) * kSBytes + ((2 * py) * kSW + 3 * px) * 4;
          uint32_t w[3][5];
#pragma unroll
          for (int dh = 0; dh < 3; ++dh)
#pragma unroll
            for (int j = 0; j < 5; ++j)
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[dh][j]) : "r"(sb + (dh * kSW + j) * 4));
          // halfword m of a row = tap (dw, ci) = m - 1; the im2col row is the 27 taps in (dh, dw, ci) order
          uint32_t pk[14];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pk[j] = __byte_perm(w[0][j], w[0][j + 1], 0x5432);
            pk[5 + j] = w[1][j + 1];
            pk[9 + j] = __byte_perm(w[2][j], w[2][j + 1], 0x5432);
          }
          pk[4] = __byte_perm(w[0][4], w[1][0], 0x7632);
          pk[13] = w[2][4] >> 16;
          // one 16-byte store per K chunk: row r's chunk sits at r*16, so a warp writes 512 contiguous
          // bytes (the 4-byte form was a 4-way bank conflict: half of the kernel's shared wavefronts)
          const uint32_t blk = a_s + (i % kRing) * kABytes + a_row;
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            const uint32_t v0 = pk[kc * 4], v1 = pk[kc * 4 + 1];
            const uint32_t v2 = kc < 3 ? pk[kc * 4 + 2] : 0u, v3 = kc < 3 ? pk[kc * 4 + 3] : 0u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + kc * (kPix * 16)), "r"(v0), "r"(v1),
                         "r"(v2), "r"(v3)
                         : "memory");
          }
          fence_proxy_async();
          mbar_arrive(&built[i % kRing]);
          issue_frame(i + 2);
        }
        if (i >= 3) epilogue(i - 3);
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
    // The kernel is bound by the latency of these 27 scattered loads per pixel and frame, so two
    // frames are kept in flight: the loads of frame i+2 are issued when frame i is published, one
    // whole iteration (publish + epilogue) before they are needed.
    uint32_t va[28], vb[28];
    auto load_frame = [&](int f, uint32_t (&vals)[28]) {
      const TI* src = src0 + f * frame_elems;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh)
#pragma unroll
        for (int e = 0; e < 9; ++e) {
          const bool ok = rv[dh] && cv[e / 3];
          vals[dh * 9 + e] = ok ? load_bf16_bits(src + dh * row_elems + e, s_lut, e % 3) : 0u;
        }
      vals[27] = 0u;
    };
    auto publish = [&](int i, uint32_t (&vals)[28]) {
      // ---- publish the im2col block of input frame i, then refill the buffer with frame i+2
      const uint32_t blk = a_s + (i % kRing) * kABytes + a_row;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = kc * 8 + j * 2;
          v[j] = k < 28 ? (vals[k] | (vals[k + 1] << 16)) : 0u;
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + kc * (kPix * 16)), "r"(v[0]), "r"(v[1]),
                     "r"(v[2]), "r"(v[3])
                     : "memory");
      }
      fence_proxy_async();
      mbar_arrive(&built[i % kRing]);
      if (i + 2 < T) load_frame(i + 2, vals);
    };
    load_frame(0, va);
    if (T > 1) load_frame(1, vb);
    for (int i = 0; i < T + 3; ++i) {
      if (i < T) {
        if (i & 1) publish(i, vb);
        else publish(i, va);
      }
      if (i >= 3) epilogue(i - 3);
    }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 4)
    const uint32_t idesc = make_idesc(kN);
    for (int t = 0; t < T; ++t) {
      const int last = (t + 2 < T) ? t + 2 : T - 1;        // newest input frame this output needs
      mbar_wait(&built[last % kRing], static_cast<uint32_t>((last / kRing) & 1));
      const int as = t & 1;
      mbar_wait(&t_empty[as], static_cast<uint32_t>(((t >> 1) & 1) ^ 1));
      tcgen05_after_sync();
      if (elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int dt = 0; dt < kKT; ++dt) {
          const int f = t - kKT / 2 + dt;
          if (f < 0 || f >= T) continue;                   // temporal zero padding
          const uint32_t a_blk = a_s + (f % kRing) * kABytes;
          const uint32_t w_blk = w_s + dt * kWBytes;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            umma_bf16(tmem_base + as * kN,
                      make_desc_nosw(a_blk + kk * 2 * (kPix * 16), kPix * 16, 128),
                      make_desc_nosw(w_blk + kk * 2 * (kN * 16), kN * 16, 128), idesc, acc);
            acc = 1;
          }
        }
        umma_commit(&t_full[as]);
      }
      __syncwarp();
    }
  }

  tcgen05_before_sync();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base) : "memory");
  }
}

constexpr size_t kSmemBytes = 128 + kRing * kABytes + kKT * kWBytes + 128 + 128 + 16 + 768 * 2 + kSRing * kSBytes;

}  // namespace stemtc
}  // namespace x3d

using namespace x3d;

template <typename TI, bool kStaged>
static int stem_tc_launch_impl(const void* in, const void* wc, const float* bias, void* out, int N, int T, int H,
                               int W, int C, const stemtc::InputNorm& nrm, cudaStream_t st) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  dim3 grid((Wo + stemtc::kTW - 1) / stemtc::kTW, (Ho + stemtc::kTH - 1) / stemtc::kTH, N);
  X3D_REQUIRE(grid.y <= 65535, X3D_ERR_UNSUPPORTED, "x3d_stem_tc_fwd: image too tall");
  static SmemOptIn optin;                           // one per template instance, per device inside
  {
    const cudaError_t e = ensure_dynamic_smem(stemtc::stem_tc_kernel<TI, kStaged>, optin, stemtc::kSmemBytes);
    X3D_REQUIRE(e == cudaSuccess, X3D_ERR_LAUNCH, "x3d_stem_tc_fwd: smem attribute: %s", cudaGetErrorString(e));
  }
  stemtc::stem_tc_kernel<TI, kStaged><<<grid, stemtc::kThreads, stemtc::kSmemBytes, st>>>(
      static_cast<const TI*>(in), static_cast<const uint4*>(wc), bias, static_cast<bf16*>(out), T, H, W, Ho, Wo,
      C, nrm);
  return check_launch("x3d_stem_tc_fwd");
}

template <typename TI>
static int stem_tc_launch(const void* in, const void* wc, const float* bias, void* out, int N, int T, int H,
                          int W, int C, const stemtc::InputNorm& nrm, cudaStream_t st) {
  if (sizeof(TI) == 2 && W % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0)
    return stem_tc_launch_impl<TI, sizeof(TI) == 2>(in, wc, bias, out, N, T, H, W, C, nrm, st);
  return stem_tc_launch_impl<TI, false>(in, wc, bias, out, N, T, H, W, C, nrm, st);
}

static int stem_tc_check(const void* in, const void* wc, const float* bias, void* out, int N, int T, int H,
                         int W, int C, int kt) {
  X3D_REQUIRE(in && wc && bias && out, X3D_ERR_INVALID_ARG, "x3d_stem_tc_fwd: null pointer");
  X3D_REQUIRE(kt == stemtc::kKT, X3D_ERR_UNSUPPORTED, "x3d_stem_tc_fwd: temporal filter %d (only 5)", kt);
  X3D_REQUIRE(C > 0 && C % 8 == 0 && C <= stemtc::kN, X3D_ERR_UNSUPPORTED, "x3d_stem_tc_fwd: C=%d (multiple of 8, <= 32)", C);
  X3D_REQUIRE(N > 0 && N <= 65535 && T > 0 && H > 0 && W > 0, X3D_ERR_INVALID_ARG, "x3d_stem_tc_fwd: bad extent");
  X3D_REQUIRE((long)H * W * 3 < (1L << 31), X3D_ERR_UNSUPPORTED, "x3d_stem_tc_fwd: frame too large");
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(wc) & 15) == 0,
              X3D_ERR_INVALID_ARG, "x3d_stem_tc_fwd: out / wc must be 16-byte aligned");
  X3D_REQUIRE(device_sm_count() > 0 && device_is_sm100(), X3D_ERR_NO_DEVICE, "x3d_stem_tc_fwd: needs an sm_100 device");
  return X3D_OK;
}

extern "C" int x3d_stem_tc_fwd(const void* in, int in_dtype, const void* wc, const float* bias,
                               void* out, int N, int T, int H, int W, int C, int kt, void* stream) {
  const int rc = stem_tc_check(in, wc, bias, out, N, T, H, W, C, kt);
  if (rc != X3D_OK) return rc;
  const stemtc::InputNorm none{};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in_dtype == X3D_BF16) return stem_tc_launch<bf16>(in, wc, bias, out, N, T, H, W, C, none, st);
  if (in_dtype == X3D_F32) return stem_tc_launch<float>(in, wc, bias, out, N, T, H, W, C, none, st);
  X3D_REQUIRE(false, X3D_ERR_INVALID_ARG, "x3d_stem_tc_fwd: in_dtype %d", in_dtype);
}

extern "C" int x3d_stem_tc_u8_fwd(const uint8_t* in, const float* mean, const float* std, float norm_value,
                                  const void* wc, const float* bias, void* out, int N, int T, int H, int W,
                                  int C, int kt, void* stream) {
  const int rc = stem_tc_check(in, wc, bias, out, N, T, H, W, C, kt);
  if (rc != X3D_OK) return rc;
  X3D_REQUIRE(mean && std, X3D_ERR_INVALID_ARG, "x3d_stem_tc_u8_fwd: null mean/std");
  X3D_REQUIRE(norm_value != 0.f && std[0] != 0.f && std[1] != 0.f && std[2] != 0.f, X3D_ERR_INVALID_ARG,
              "x3d_stem_tc_u8_fwd: zero divisor");
  stemtc::InputNorm nrm;
  for (int i = 0; i < 3; ++i) { nrm.mean[i] = mean[i]; nrm.std[i] = std[i]; }
  nrm.norm_value = norm_value;
  return stem_tc_launch<uint8_t>(in, wc, bias, out, N, T, H, W, C, nrm, static_cast<cudaStream_t>(stream));
}
