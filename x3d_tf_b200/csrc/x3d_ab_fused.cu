// Fused bottleneck front half:  1x1x1 expand conv + BN + ReLU  ->  channelwise 3x3x3 conv + BN
// (+ SE partial sums), i.e. Bottleneck.a/bn_a/relu/b/bn_b (+ the reduction of se_pool), reference
// model.py:306-312, as ONE kernel.  The `inner`-wide tensor between the two convolutions -- the
// widest tensor of the network, 30 % + 35 % of all HBM bytes in a layer-by-layer schedule
//  -- never leaves the SM.
//
// One CTA = one clip x one chunk of CH inner channels x one spatial tile of Q x SW outputs; it
// marches over the T frames like the unfused channelwise kernel (x3d_dw_tma.cu).  Per frame t:
//   1. the copy/MMA lane TMA-loads the halo tile of the block INPUT (BH x BW pixels x <=64*KC
//      channels, 128B swizzle; out-of-image pixels and channels are zero-filled by the TMA unit)
//      into a 2-deep A ring;
//   2. the same lane issues tcgen05.mma (M=128 per pixel block, N=CHN, K=16) against the CTA's
//      resident, BN-folded weight slice Wa[c0:c0+CHN, :]; the fp32 result lands in TMEM;
//   3. the compute warps drain TMEM (tcgen05.ld; warp w owns lane quarter w%4), add the BN
//      shift, apply ReLU, force pixels outside the image to zero (TF 'SAME' pads the OUTPUT of
//      `a`, which is relu(shift) != 0 where the input is padding) and write bf16 into a 2-deep
//      shared-memory frame ring;
//   4. the same warps run the 27-tap stencil over that frame exactly as x3d_dw_tma.cu does
//      (one channel pair x one output row per thread, 9 packed FFMA2 per staged value, three
//      rotating accumulator sets), pack the finished output frame t-1 into a staging buffer and
//      the copy lane TMA-stores it.
// Hand-over between the roles is by mbarriers only; the frame loop has no __syncthreads.
#include <stdlib.h>

#include "tma_common.cuh"

namespace x3d {
namespace abf {

using namespace ptx;

constexpr int kRing = 2;      // A ring and output staging depth (index math below assumes 2)
constexpr int kFrames = 3;    // frame ring depth: frame t+1 is drained while frame t is read
constexpr int kStoreLag = 3;  // the copy lane stores frame t-3 in iteration t (never waits on the step in flight)

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
// K-major operand, 128-byte swizzle, 8-row atoms stacked every 1024 B (same as x3d_pw_tc.cu).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t sbo = 1024) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(128 >> 4) << 24);
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t dst, const CUtensorMap* map, int c0, int c1,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct Params {
  const float* bias_a;   // [Cs] BN shift of bn_a
  const float* w;        // [27, Cs] BN-folded channelwise taps
  const float* bias;     // [Cs] BN shift of bn_b
  float* partial;        // [N, tiles, Cs] or nullptr
  int T, H, W, Ho, Wo, Cs;
  int Q, tiles_w, tiles;
  int pad_h, pad_w;
  int cwarps;            // compute warps; warp `cwarps` is the copy/MMA warp
  int KC;                // 64-wide K chunks of the expand GEMM
  int k16;               // K=16 MMA steps in total (= ceil(Cin/16))
  int MB;                // 128-pixel blocks of the halo tile
  int CHN;               // MMA N (CH rounded up to 16)
  int a_kc_bytes;        // one K chunk of an A stage: halo pixels rounded up to 8 rows x 128 B
  int a_stage_bytes;     // KC * a_kc_bytes
  int a_box_bytes;       // bytes one x-tile TMA box delivers (per K chunk)
  int slot_bytes;        // frame ring slot
  int stage_bytes;       // output staging buffer
  int off_wa, off_a, off_ring, off_stage, off_red, off_w2, off_bias;   // from the 1024-aligned base
  int dbg;               // X3D_ABF_DEBUG experiments (0 in production)
};

template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float2 lds2(uint32_t a) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
    return r;
  }
};
template <> struct Elem<bf16> {
  static __device__ __forceinline__ float2 lds2(uint32_t a) {
    // byte permutes keep the unpack on the ALU pipe (ptxas turns `u << 16` into an IMAD, which
    // would compete with the FFMA2 stream for the FMA pipe)
    uint32_t u, lo, hi;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(a));
    asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(u));
    asm("prmt.b32 %0, %1, 0, 0x3244;" : "=r"(hi) : "r"(u));
    return make_float2(__uint_as_float(lo), __uint_as_float(hi));
  }
};
__device__ __forceinline__ void sts2_f32(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts2_bf16(uint32_t a, float2 v) {
  __nv_bfloat162 h = __float22bfloat162_rn(v);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(*reinterpret_cast<uint32_t*>(&h)) : "memory");
}

// Pixel pitch of the frame ring: CH bf16, padded so that pitch/16 is odd (32 lanes = 32
// consecutive pixels then write their 16-byte vectors conflict-free).
template <int CH> struct RingPitch { static constexpr int value = ((CH * 2 / 16) & 1) ? CH * 2 : CH * 2 + 16; };

template <int S, int SW, int CH>
__global__ void __launch_bounds__(256, 2)
ab_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmOut, const Params p) {
  constexpr int BW = (SW - 1) * S + 3;
  constexpr int PS = RingPitch<CH>::value;   // bytes per ring pixel
  constexpr int RS = BW * PS;                // bytes per ring row
  constexpr int OPS = CH * 2;                // bytes per staged output pixel (dense: TMA store)
  constexpr int C2 = CH / 2;

  extern __shared__ __align__(1024) uint8_t ab_smem_raw[];
  const uint32_t raw_s = smem_u32(ab_smem_raw);
  const uint32_t smem_s = (raw_s + 1023u) & ~1023u;
  uint8_t* smem = ab_smem_raw + (smem_s - raw_s);
  // barriers live in the first 256 bytes
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);     // weights landed
  uint64_t* a_full = w_full + 1;                            // [2] x tile landed
  uint64_t* a_empty = a_full + kRing;                       // [2] MMAs reading the stage retired
  uint64_t* acc_full = a_empty + kRing;                     // [2] TMEM accumulator complete
  uint64_t* acc_empty = acc_full + 2;                       // [2] TMEM accumulator drained
  uint64_t* ring_full = acc_empty + 2;                      // [3] frame written by every warp
  uint64_t* ring_empty = ring_full + kFrames;               // [3] frame read by every warp
  uint64_t* staged = ring_empty + kFrames;                  // [2] output frame packed
  uint64_t* sfree = staged + kRing;                         // [2] TMA store has read the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfree + kRing);
  const uint32_t wa_s = smem_s + p.off_wa;
  const uint32_t a_s = smem_s + p.off_a;
  const uint32_t ring_s = smem_s + p.off_ring;
  const uint32_t stage_s = smem_s + p.off_stage;
  float* s_red = reinterpret_cast<float*>(smem + p.off_red);
  const uint32_t w2_s = smem_s + p.off_w2;
  // bias as a GEMM operand: ones[128 x 16] (one 8-row atom, SBO = 0) x  [hi(shift), lo(shift)] rows
  const uint32_t ones_s = smem_s + p.off_bias;               // 1 KiB atom
  const uint32_t biasop_s = ones_s + 1024;                   // [CHN rows x 128 B], 128B swizzle

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int slot = tid / C2, cp = tid - slot * C2;
  const int n = blockIdx.z, c0 = blockIdx.y * CH;
  const int tile_h = blockIdx.x / p.tiles_w, tile_w = blockIdx.x - tile_h * p.tiles_w;
  const int ho0 = tile_h * p.Q, wo0 = tile_w * SW;
  const int c = c0 + 2 * cp;
  const bool is_copy = warp == p.cwarps;
  const bool in_slot = !is_copy && slot < p.Q;
  const bool on = in_slot && ho0 + slot < p.Ho && c < p.Cs;
  const int hi0 = ho0 * S - p.pad_h, wi0 = wo0 * S - p.pad_w;
  const int BH = (p.Q - 1) * S + 3;
  const int n_acc = p.MB == 1 ? 2 : 1;                      // TMEM accumulator buffers

  // The BN shift of bn_a enters the accumulator through the tensor core as well: D = ones x Bop
  // with ones[m, 0:2] = 1 and Bop[n, 0:2] = (hi, lo) bf16 split of shift[n] (exact to 2^-17), so
  // the drain does no floating-point add at all.  Both tiles are K-major, 128B-swizzled: the
  // 16-byte chunk j of row r sits at chunk position j ^ (r % 8).
  for (int i = tid; i < 8 + p.CHN; i += blockDim.x) {
    const bool is_one = i < 8;
    const int r = is_one ? i : i - 8;
    const uint32_t row_s = (is_one ? ones_s : biasop_s) + static_cast<uint32_t>(r) * 128u;
    uint32_t w0 = 0x3f803f80u;                                 // (1.0, 1.0) bf16
    if (!is_one) {
      const float b = (c0 + r < p.Cs) ? p.bias_a[c0 + r] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(b);
      const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
      w0 = static_cast<uint32_t>(__bfloat16_as_ushort(hi)) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo)) << 16);
    }
    for (int j = 0; j < 8; ++j) {
      const uint32_t a = row_s + static_cast<uint32_t>((j ^ (r & 7)) * 16);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(j == 0 ? w0 : 0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
    }
  }
  fence_proxy_async();                                         // generic writes -> tensor-core (async proxy) reads
  if (tid == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmOut);
    mbar_init(w_full, 1);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], static_cast<uint32_t>(p.cwarps));
      mbar_init(&staged[s], static_cast<uint32_t>(p.cwarps));
      mbar_init(&sfree[s], 1);
    }
    for (int s = 0; s < kFrames; ++s) {
      mbar_init(&ring_full[s], static_cast<uint32_t>(p.cwarps));
      mbar_init(&ring_empty[s], static_cast<uint32_t>(p.cwarps));
    }
    fence_barrier_init();
  }
  if (is_copy) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (is_copy) {
    // ------------------------------------------------------------ copy / MMA warp (one lane)
    if (lane == 0) {
      const int w_chunk_bytes = p.CHN * 128;
      mbar_expect_tx(w_full, static_cast<uint32_t>(p.KC * w_chunk_bytes));
      for (int kc = 0; kc < p.KC; ++kc)
        tma_load_2d_s(wa_s + kc * w_chunk_bytes, &tmW, kc * 64, c0, w_full);
      auto load_x = [&](int f) {
        const int s = f & 1;
        mbar_expect_tx(&a_full[s], static_cast<uint32_t>(p.KC * p.a_box_bytes));
        for (int kc = 0; kc < p.KC; ++kc)
          tma_load_5d(a_s + s * p.a_stage_bytes + kc * p.a_kc_bytes, &tmX, kc * 64, wi0, hi0, f, n,
                      &a_full[s]);
      };
      for (int f = 0; f < kRing && f < p.T; ++f) load_x(f);
      const uint32_t idesc = make_idesc_bf16(p.CHN);
      mbar_wait(w_full, 0);
      auto store_frame = [&](int f) {
        const int k = f & 1;
        mbar_wait_lean(&staged[k], static_cast<uint32_t>((f >> 1) & 1));
        tma_store_5d(&tmOut, stage_s + k * p.stage_bytes, c0, wo0, ho0, f, n);
        tma_store_commit();
        // never wait for the store just issued: only for the previous one (a whole step old), whose
        // staging buffer is the one the compute warps pack next
        tma_store_wait_read<1>();
        if (f >= 1) mbar_arrive(&sfree[(f - 1) & 1]);
      };
      for (int t = 0; t < p.T; ++t) {
        const int s = t & 1;
        const int b = n_acc == 2 ? (t & 1) : 0;
        mbar_wait_lean(&a_full[s], static_cast<uint32_t>((t >> 1) & 1));
        mbar_wait_lean(&acc_empty[b], static_cast<uint32_t>((n_acc == 2 ? (t >> 1) & 1 : t & 1) ^ 1));
        tcgen05_after_sync();
        const uint32_t a_base = a_s + s * p.a_stage_bytes;
        for (int mb = 0; mb < ((p.dbg & 2) ? 0 : p.MB); ++mb) {
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b * 128 + mb * p.CHN);
          umma_bf16(d_tmem, make_desc_sw128(ones_s, 0), make_desc_sw128(biasop_s), idesc, 0u);
          for (int k = 0; k < p.k16; ++k) {
            const int kc = k >> 2, kk = k & 3;
            umma_bf16(d_tmem, make_desc_sw128(a_base + kc * p.a_kc_bytes + mb * 16384 + kk * 32),
                      make_desc_sw128(wa_s + kc * w_chunk_bytes + kk * 32), idesc, 1u);
          }
        }
        umma_commit(&a_empty[s]);
        umma_commit(&acc_full[b]);
        if (t + kRing < p.T) {               // refill the A stage as soon as its MMAs have retired
          mbar_wait_lean(&a_empty[s], static_cast<uint32_t>((t >> 1) & 1));
          load_x(t + kRing);
        }
        if (t >= kStoreLag) store_frame(t - kStoreLag);
      }
      for (int f = (p.T > kStoreLag ? p.T - kStoreLag : 0); f < p.T; ++f) store_frame(f);
      tma_store_wait_read<0>();              // shared memory must outlive the bulk stores
    }
  } else {
    // ------------------------------------------------------------ compute warps
    // dt=0 taps stay in registers; the dt=1 and dt=2 taps are read from shared memory ([18][C2]
    // float2, conflict-free) so that the TMEM drain fits in the 128-register budget without spills
    float2 wr[9];
    float2 bia = make_float2(0.f, 0.f);
    if (on) {
#pragma unroll
      for (int i = 0; i < 9; ++i) wr[i] = ld2(p.w + i * p.Cs + c);
      bia = ld2(p.bias + c);
    } else {
#pragma unroll
      for (int i = 0; i < 9; ++i) wr[i] = make_float2(0.f, 0.f);
    }
    const uint32_t w2_t = w2_s + static_cast<uint32_t>(cp) * 8;
    if (slot == 0) {
#pragma unroll
      for (int i = 0; i < 18; ++i) {
        const float2 w = (c < p.Cs) ? ld2(p.w + (9 + i) * p.Cs + c) : make_float2(0.f, 0.f);
        sts2_f32(w2_t + i * (C2 * 8), w);
      }
    }
    const uint32_t cthreads = static_cast<uint32_t>(p.cwarps) * 32u;
    asm volatile("bar.sync 1, %0;" ::"r"(cthreads) : "memory");

    float2 acc[3][SW];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int j = 0; j < SW; ++j) acc[a][j] = bia;
    float2 ssum = make_float2(0.f, 0.f);

    const int aslot = in_slot ? slot : 0;
    const uint32_t toff = static_cast<uint32_t>(aslot * S) * RS + static_cast<uint32_t>(cp) * 4;
    const uint32_t soff = static_cast<uint32_t>(aslot * SW) * OPS + static_cast<uint32_t>(cp) * 4;
    int ncol = p.Wo - wo0;
    if (ncol > SW) ncol = SW;
    const bool lane0 = lane == 0;
    const bool on_se = on && p.partial != nullptr;

    // ---- TMEM drain assignment: lane quarter q = warp % 4; the warps that share a quarter split
    // the CHN columns in groups of 8.
    const int q = warp & 3;
    const int peers = (p.cwarps - 1 - q) / 4 + 1;
    const int me = warp >> 2;
    const int groups = (CH + 7) / 8;                         // 8-column groups that hold real channels
    const int g_lo = groups * me / peers, g_hi = groups * (me + 1) / peers;
    const int npix = BH * BW;

    // per-thread pixel bookkeeping of the drain, once per CTA: for pixel block mb this lane owns halo
    // pixel r = mb*128 + q*32 + lane; bit mb of pix_m = it exists, of in_m = it is inside the image
    uint32_t pix_m = 0, in_m = 0;
    int mb_n = 0;                                              // pixel blocks this lane quarter touches
    for (int mb = 0; mb < p.MB; ++mb) {
      if (mb * 128 + q * 32 >= npix) break;
      mb_n = mb + 1;
      const int r = mb * 128 + q * 32 + lane;
      const int hh = r / BW, ww = r - hh * BW;
      if (r < npix) {
        pix_m |= 1u << mb;
        if (static_cast<unsigned>(hi0 + hh) < static_cast<unsigned>(p.H) &&
            static_cast<unsigned>(wi0 + ww) < static_cast<unsigned>(p.W))
          in_m |= 1u << mb;
      }
    }
    const uint32_t drain_off = static_cast<uint32_t>(q * 32 + lane) * PS + static_cast<uint32_t>(g_lo) * 16;
    const uint32_t drain_tm = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g_lo * 8);

    auto emit8 = [&](uint32_t dst, bool inside, bool pix, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                     uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7) {
      uint32_t o0, o1, o2, o3;
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o0) : "f"(__uint_as_float(a1)), "f"(__uint_as_float(a0)));
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o1) : "f"(__uint_as_float(a3)), "f"(__uint_as_float(a2)));
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o2) : "f"(__uint_as_float(a5)), "f"(__uint_as_float(a4)));
      asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(o3) : "f"(__uint_as_float(a7)), "f"(__uint_as_float(a6)));
      if (inside)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
      else if (pix)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
    };

    // Index arithmetic is shifts and masks only (IMAD-based div/mod would sit on the FMA pipe the
    // stencil saturates): frame t = 3*k3 + rs with rs known at compile time in the unrolled loop.
    auto drain = [&](int t, int rs, int k3) {
      const int b = n_acc == 2 ? (t & 1) : 0;
      const uint32_t acc_par = static_cast<uint32_t>(n_acc == 2 ? (t >> 1) & 1 : t & 1);
      mbar_wait_lean(&acc_full[b], acc_par);
      tcgen05_after_sync();
      if (k3 >= 1) mbar_wait_lean(&ring_empty[rs], static_cast<uint32_t>((k3 & 1) ^ 1));
      const uint32_t slot_base = ring_s + rs * p.slot_bytes + drain_off;
      for (int mb = 0; mb < ((p.dbg & 1) ? 0 : mb_n); ++mb) {
        const bool pix = (pix_m >> mb) & 1u, inside = (in_m >> mb) & 1u;
        const uint32_t taddr = drain_tm + static_cast<uint32_t>(b * 128 + mb * p.CHN);
        const uint32_t dst = slot_base + static_cast<uint32_t>(mb * 128) * PS;
        int g = g_lo;
        for (; g + 1 < g_hi; g += 2) {
          uint32_t v[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                "=r"(v[14]), "=r"(v[15])
              : "r"(taddr + (g - g_lo) * 8));
          tmem_ld_wait();
          emit8(dst + (g - g_lo) * 16, inside, pix, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
          emit8(dst + (g - g_lo) * 16 + 16, inside, pix, v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
        }
        if (g < g_hi) {
          uint32_t v[8];
          tmem_ld8(taddr + (g - g_lo) * 8, v);
          tmem_ld_wait();
          emit8(dst + (g - g_lo) * 16, inside, pix, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
        }
      }
      tcgen05_before_sync();
      __syncwarp();
      if (lane0) {
        mbar_arrive(&acc_empty[b]);
        mbar_arrive(&ring_full[rs]);
      }
    };

    auto stage_out = [&](float2 (&A)[SW], int t_out) {
      const int k = t_out & 1;
      if (t_out >= kRing) mbar_wait_lean(&sfree[k], static_cast<uint32_t>(((t_out >> 1) & 1) ^ 1));
      const uint32_t dst = stage_s + k * p.stage_bytes + soff;
      if (in_slot) {
#pragma unroll
        for (int j = 0; j < SW; ++j) sts2_bf16(dst + j * OPS, A[j]);
      }
      if (on_se) {
#pragma unroll
        for (int j = 0; j < SW; ++j)
          if (j < ncol) ssum = __fadd2_rn(ssum, A[j]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane0) mbar_arrive(&staged[k]);
    };

    auto step = [&](int t, int rs, int k3, float2 (&A0)[SW], float2 (&A1)[SW], float2 (&A2)[SW]) {
      // next frame first: its hand-over is never waited for
      if (t + 1 < p.T) drain(t + 1, rs == 2 ? 0 : rs + 1, rs == 2 ? k3 + 1 : k3);
      mbar_wait_lean(&ring_full[rs], static_cast<uint32_t>(k3 & 1));
      const uint32_t base = ring_s + rs * p.slot_bytes + toff;
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        float2 w1[3], w2[3];
#pragma unroll
        for (int dw = 0; dw < 3; ++dw) {
          w1[dw] = Elem<float>::lds2(w2_t + (dh * 3 + dw) * (C2 * 8));
          w2[dw] = Elem<float>::lds2(w2_t + (9 + dh * 3 + dw) * (C2 * 8));
        }
#pragma unroll
        for (int jj = 0; jj < BW; ++jj) {
          const float2 x = Elem<bf16>::lds2(base + dh * RS + jj * PS);
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const int jn = jj - dw;
            if (jn >= 0 && jn % S == 0 && jn / S < SW) {
              const int j = jn / S;
              A0[j] = fma2(x, wr[dh * 3 + dw], (dh == 0 && dw == 0) ? bia : A0[j]);
              A1[j] = fma2(x, w1[dw], A1[j]);
              A2[j] = fma2(x, w2[dw], A2[j]);
            }
          }
        }
      }
      __syncwarp();
      if (lane0) mbar_arrive(&ring_empty[rs]);
      if (t >= 1) stage_out(A2, t - 1);
    };

    drain(0, 0, 0);
    for (int t = 0, k3 = 0; t < p.T; t += 3, ++k3) {
      step(t, 0, k3, acc[1], acc[0], acc[2]);
      if (t + 1 < p.T) step(t + 1, 1, k3, acc[2], acc[1], acc[0]);
      if (t + 2 < p.T) step(t + 2, 2, k3, acc[0], acc[2], acc[1]);
    }
    {
      const int r = (p.T - 1) % 3;
      if (r == 0) stage_out(acc[0], p.T - 1);
      else if (r == 1) stage_out(acc[1], p.T - 1);
      else stage_out(acc[2], p.T - 1);
    }
    if (p.partial != nullptr) {
      if (in_slot) {
        s_red[slot * CH + 2 * cp] = ssum.x;
        s_red[slot * CH + 2 * cp + 1] = ssum.y;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(cthreads) : "memory");
      for (int ch = tid; ch < CH; ch += static_cast<int>(cthreads)) {
        if (c0 + ch < p.Cs) {
          float a = 0.f;
          for (int k = 0; k < p.Q; ++k) a += s_red[k * CH + ch];
          p.partial[(static_cast<long>(n) * p.tiles + blockIdx.x) * p.Cs + c0 + ch] = a;
        }
      }
    }
  }

  tcgen05_before_sync();
  __syncthreads();
  if (is_copy) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------ host
struct Plan {
  int CH, SW, Q, cwarps, threads, tiles_w, tiles_h, chunks, BH, BW, MB, KC, k16, CHN;
  int a_kc_bytes, a_stage_bytes, a_box_bytes, slot_bytes, stage_bytes;
  int off_wa, off_a, off_ring, off_stage, off_red, off_w2, off_bias;
  size_t smem;
  bool ok;
};

static int ring_pitch(int CH) { return ((CH * 2 / 16) & 1) ? CH * 2 : CH * 2 + 16; }

// Same objective as the unfused kernel (padded output volume first), restricted to tiles whose
// halo fits 3 pixel blocks of 128, >= 4 compute warps (every TMEM lane quarter needs an owner) and,
// preferably, two CTAs per SM.
static Plan make_plan(int H, int W, int Cin, int Cs, int stride, int max_smem) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const int chs[3] = {56, 64, 72};
  const int sws1[2] = {8, 7}, sws2[3] = {8, 7, 4};
  const int nsw = stride == 1 ? 2 : 3;
  Plan best{};
  best.ok = false;
  double best_cost = 1e300;
  for (int ci = 0; ci < 3; ++ci) {
    for (int si = 0; si < nsw; ++si) {
      const int CH = chs[ci], SW = stride == 1 ? sws1[si] : sws2[si];
      int qmax = 224 / (CH / 2);                     // <= 7 compute warps + the copy warp
      if (qmax > 9) qmax = 9;
      if (qmax > Ho) qmax = Ho;
      for (int Q = qmax; Q >= 1; --Q) {
        Plan pl{};
        pl.CH = CH; pl.SW = SW; pl.Q = Q;
        pl.chunks = (Cs + CH - 1) / CH;
        pl.BW = (SW - 1) * stride + 3;
        pl.BH = (Q - 1) * stride + 3;
        pl.MB = (pl.BH * pl.BW + 127) / 128;
        if (pl.MB > 3) continue;
        pl.CHN = (CH + 15) / 16 * 16;
        if (pl.MB * pl.CHN > 256) continue;
        pl.k16 = (Cin + 15) / 16;
        pl.KC = (pl.k16 + 3) / 4;
        pl.cwarps = (Q * (CH / 2) + 31) / 32;
        if (pl.cwarps < 4) continue;
        pl.threads = pl.cwarps * 32 + 32;
        pl.a_kc_bytes = (pl.BH * pl.BW + 7) / 8 * 1024;
        pl.a_stage_bytes = pl.KC * pl.a_kc_bytes;
        pl.a_box_bytes = pl.BH * pl.BW * 128;
        pl.slot_bytes = (pl.BH * pl.BW * ring_pitch(CH) + 127) / 128 * 128;
        pl.stage_bytes = (Q * SW * CH * 2 + 127) / 128 * 128;
        int off = 256;                                   // barriers + tmem slot
        off = (off + 1023) / 1024 * 1024;
        pl.off_wa = off;   off += pl.KC * pl.CHN * 128;
        off = (off + 1023) / 1024 * 1024;
        pl.off_a = off;    off += kRing * pl.a_stage_bytes;
        pl.off_ring = off; off += kFrames * pl.slot_bytes;
        pl.off_stage = off; off += kRing * pl.stage_bytes;
        pl.off_red = off;  off += 9 * CH * 4;
        pl.off_w2 = off;   off += 27 * CH * 4;
        off = (off + 1023) / 1024 * 1024;
        pl.off_bias = off; off += 1024 + pl.CHN * 128;   // ones atom + bias operand tile
        // the last pixel block of an MMA always reads 128 rows: keep that inside the allocation
        const int mma_end = pl.off_a + kRing * pl.a_stage_bytes - pl.a_kc_bytes + pl.MB * 16384;
        if (off < mma_end) off = mma_end;
        pl.smem = (size_t)off + 1024;                    // + alignment slack
        if ((int)pl.smem > max_smem) continue;
        pl.tiles_w = (Wo + SW - 1) / SW;
        pl.tiles_h = (Ho + Q - 1) / Q;
        const double tiles = (double)pl.tiles_w * pl.tiles_h * pl.chunks;
        const double work = tiles * Q * SW * CH;
        const double staged = tiles * pl.BH * pl.BW * CH;
        double cost = work + 0.15 * staged + 1e-3 * tiles;
        if (pl.smem > 113 * 1024) cost *= 1.35;          // one CTA per SM only
        pl.ok = true;
        if (cost < best_cost) { best_cost = cost; best = pl; }
      }
    }
  }
  return best;
}

template <int S, int SW, int CH>
static int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& to, const Params& p,
                  const Plan& pl, int N, cudaStream_t st) {
  auto kern = ab_fused_kernel<S, SW, CH>;
  static SmemOptIn optin;                              // one per template instance, per device inside
  const cudaError_t e = ensure_dynamic_smem(kern, optin, pl.smem);
  if (e != cudaSuccess) {
    set_error("x3d_expand_dw_fwd: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e));
    return X3D_ERR_LAUNCH;
  }
  dim3 grid(pl.tiles_w * pl.tiles_h, pl.chunks, N);
  kern<<<grid, pl.threads, pl.smem, st>>>(tx, tw, to, p);
  return check_launch("x3d_expand_dw_fwd");
}

template <int S>
static int dispatch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& to, const Params& p,
                    const Plan& pl, int N, cudaStream_t st) {
#define X3D_ABF(SWW, CHH) \
  if (pl.SW == SWW && pl.CH == CHH) return launch<S, SWW, CHH>(tx, tw, to, p, pl, N, st)
  X3D_ABF(8, 56); X3D_ABF(8, 64); X3D_ABF(8, 72);
  X3D_ABF(7, 56); X3D_ABF(7, 64); X3D_ABF(7, 72);
  if constexpr (S == 2) { X3D_ABF(4, 56); X3D_ABF(4, 64); X3D_ABF(4, 72); }
#undef X3D_ABF
  set_error("x3d_expand_dw_fwd: no kernel for SW=%d CH=%d", pl.SW, pl.CH);
  return X3D_ERR_UNSUPPORTED;
}

}  // namespace abf
}  // namespace x3d

using namespace x3d;

extern "C" int x3d_expand_dw_partial_blocks(int T, int H, int W, int Cin, int C, int stride) {
  if (T <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || Cin <= 0 || Cin % 8 || (stride != 1 && stride != 2)) return 0;
  const int ms = device_max_smem();
  if (ms <= 0) return 0;
  const abf::Plan pl = abf::make_plan(H, W, Cin, C, stride, ms);
  return pl.ok ? pl.tiles_w * pl.tiles_h : 0;
}

extern "C" int x3d_expand_dw_fwd(const void* x, const void* wa, const float* bias_a, const float* wb,
                                 const float* bias_b, void* out, float* se_partial, int N, int T, int H,
                                 int W, int Cin, int C, int Kpad, int Npad, int stride, int pad_h,
                                 int pad_w, void* stream) {
  X3D_REQUIRE(x && wa && bias_a && wb && bias_b && out, X3D_ERR_INVALID_ARG, "x3d_expand_dw_fwd: null pointer");
  X3D_REQUIRE(C > 0 && C % 8 == 0 && Cin > 0 && Cin % 8 == 0, X3D_ERR_INVALID_ARG,
              "x3d_expand_dw_fwd: Cin=%d / C=%d must be multiples of 8", Cin, C);
  X3D_REQUIRE(Kpad % 64 == 0 && Kpad >= Cin && Npad % 16 == 0 && Npad >= C, X3D_ERR_INVALID_ARG,
              "x3d_expand_dw_fwd: bad packed weight extents Kpad=%d Npad=%d", Kpad, Npad);
  X3D_REQUIRE(stride == 1 || stride == 2, X3D_ERR_UNSUPPORTED, "x3d_expand_dw_fwd: stride %d", stride);
  X3D_REQUIRE(N > 0 && N <= 65535 && T > 0 && H > 0 && W > 0, X3D_ERR_INVALID_ARG, "x3d_expand_dw_fwd: bad extent");
  X3D_REQUIRE(pad_h >= 0 && pad_h <= 1 && pad_w >= 0 && pad_w <= 1, X3D_ERR_INVALID_ARG, "x3d_expand_dw_fwd: pad_before must be 0 or 1");
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(wa) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_expand_dw_fwd: pointers must be 16-byte aligned");
  X3D_REQUIRE(device_sm_count() > 0 && device_is_sm100(), X3D_ERR_NO_DEVICE, "x3d_expand_dw_fwd: needs an sm_100 device");
  EncodeTiledFn enc = tensor_map_encoder();
  X3D_REQUIRE(enc != nullptr, X3D_ERR_NO_DEVICE, "x3d_expand_dw_fwd: cuTensorMapEncodeTiled unavailable");
  const abf::Plan pl = abf::make_plan(H, W, Cin, C, stride, device_max_smem());
  X3D_REQUIRE(pl.ok, X3D_ERR_UNSUPPORTED, "x3d_expand_dw_fwd: no tile plan for H=%d W=%d Cin=%d C=%d stride=%d", H, W, Cin, C, stride);
  X3D_REQUIRE(pl.chunks <= 65535, X3D_ERR_UNSUPPORTED, "x3d_expand_dw_fwd: too many channel chunks");

  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMap tx, tw, to;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                             (cuuint64_t)T * H * W * Cin * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, 1, 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_expand_dw_fwd: input tensor map failed (%d) for [%d,%d,%d,%d,%d] box [64,%d,%d]",
                (int)r, N, T, H, W, Cin, pl.BW, pl.BH);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Npad};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)pl.CHN};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wa), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_expand_dw_fwd: weight tensor map failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)Wo * C * 2, (cuuint64_t)Ho * Wo * C * 2,
                             (cuuint64_t)T * Ho * Wo * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)pl.CH, (cuuint32_t)pl.SW, (cuuint32_t)pl.Q, 1, 1};
    CUresult r = enc(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_expand_dw_fwd: output tensor map failed (%d)", (int)r);
  }

  abf::Params p;
  p.bias_a = bias_a; p.w = wb; p.bias = bias_b; p.partial = se_partial;
  p.T = T; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.Cs = C;
  p.Q = pl.Q; p.tiles_w = pl.tiles_w; p.tiles = pl.tiles_w * pl.tiles_h;
  p.pad_h = pad_h; p.pad_w = pad_w; p.cwarps = pl.cwarps;
  p.KC = pl.KC; p.k16 = pl.k16; p.MB = pl.MB; p.CHN = pl.CHN;
  p.a_kc_bytes = pl.a_kc_bytes; p.a_stage_bytes = pl.a_stage_bytes; p.a_box_bytes = pl.a_box_bytes;
  p.slot_bytes = pl.slot_bytes; p.stage_bytes = pl.stage_bytes;
  p.off_wa = pl.off_wa; p.off_a = pl.off_a; p.off_ring = pl.off_ring; p.off_stage = pl.off_stage;
  p.off_red = pl.off_red; p.off_w2 = pl.off_w2; p.off_bias = pl.off_bias;
  { const char* e = getenv("X3D_ABF_DEBUG"); p.dbg = e ? atoi(e) : 0; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return stride == 1 ? abf::dispatch<1>(tx, tw, to, p, pl, N, st) : abf::dispatch<2>(tx, tw, to, p, pl, N, st);
}
