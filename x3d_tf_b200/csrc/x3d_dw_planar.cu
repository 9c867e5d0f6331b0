// Channelwise 3x3x3 convolution + folded BN (+ swish) (+ SE partial sums), "planar" form:
// lanes = pixels, warp = channel pair, tap weights in UNIFORM registers.
// Replaces Bottleneck.b + bn_b (+ the reduction of se_pool, + the swish of blocks without SE),
// reference model.py:309-316 -- same contract as x3d_dw3x3x3_act_fwd (x3d_dw_tma.cu).
//
// Why a second stencil kernel: in x3d_dw_tma.cu a thread owns a channel pair, so every packed FFMA2
// reads three 64-bit register pairs (x, tap, accumulator).  Without an operand-cache hit that is
// three register-file cycles for a two-cycle pipe slot, and together with the bf16 unpack and the
// tap loads the kernel stays at ~49 % of the FMA pipe.  Here the 32 lanes of a warp are 32 PIXELS
// of one channel pair: the 27 taps are the same for the whole warp, ptxas keeps them in uniform
// registers (the pair index is made provably warp-uniform with a shuffle) and the stencil's
// instruction is  FFMA2 Racc, Rx, URtap, Racc  -- two register operands, no reuse needed.  The
// price is a layout change on the way in: pixels of one channel pair must be contiguous in shared
// memory, NDHWC has the channels contiguous.  So:
//   warp 0  lane 0   TMA producer: halo tile of the input for (item, frame), 16 channels wide
//                    ([BH][BW][16] bf16, zero-filled outside the tensor = TF 'SAME' padding);
//   warp 1           output converter: planar bf16x2 frame [8 pairs][TH*LW] (written by the stencil
//                    warps with conflict-free 4-byte stores) -> dense NDHWC tile [TH][LW][16]
//                    (4 x LDS.32 + STS.128 per 8 channels), then one lane issues the TMA store;
//   warps 2-7        transposers: raw tile -> fp32 planar ring [8 pairs][BH][BW] (LDS.128, unpack,
//                    STS.64): the bf16 unpack happens once per value here, not once per use;
//   warps 8-15       stencil: warp = channel pair, lane = (column, row strip), Q output rows per
//                    thread; a ring value is read with one conflict-free LDS.64 and scattered into
//                    up to 3 output rows x 3 output frames.
// One CTA = 512 threads (setmaxnreg: helpers 40, stencil 88 registers), two CTAs per SM, persistent over (clip, tile, 16-channel chunk) items;
// hand-over between the roles by mbarriers only.
#include <stdlib.h>

#include "tma_common.cuh"

namespace x3d {
namespace dwp {

using namespace ptx;

constexpr int kCh = 16, kPairs = kCh / 2;            // channels / channel pairs per item
constexpr int kXfWarp0 = 2, kXfWarps = 6;
constexpr int kStWarp0 = 8, kStWarps = kPairs;
constexpr int kThreads = 32 * (kStWarp0 + kStWarps);   // 512: warpgroups 0-1 = helpers, 2-3 = stencil
// 2 CTAs x 512 threads start with 64 registers per thread; the helper warpgroups give 24 of theirs
// back and the stencil warpgroups take them (8 x 40 + 8 x 88 = 16 x 64)
constexpr int kRegsHelper = 40, kRegsStencil = 88;

template <int N> __device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
constexpr int kRaw = 2, kRing = 2, kOut = 2;         // ring depths: raw tiles, planar frames, output staging
constexpr int kMaxPairs = 288;                       // constant-bank tap table: 288 pairs x 28 float2 = 63 KiB

// [pair][27 taps + shift] as (channel 2p, channel 2p+1); refreshed by a device-to-device copy before
// every launch (stream-ordered; the table is per layer)
__constant__ float2 c_taps[kMaxPairs * 28];

__device__ __forceinline__ float2 lds2_f32(uint32_t a) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts2_f32(uint32_t a, float lo, float hi) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(lo), "f"(hi) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// Barrier wait: hardware-suspended try_wait (the thread sleeps until the barrier's phase completes or the
// time hint expires), retried; a lost arrival traps instead of hanging.  An earlier form slept with
// `nanosleep` between non-blocking tests to save the issue slots of the polling helper warps: equal speed on
// one box, but THREE TIMES slower on another B200 box of the same pool (stage-2 layer 0.45 -> 1.34 ms), where
// the sleep evidently lasts far longer than asked -- nanosleep's duration is only bounded by 2x the argument
// on top of the timer's resolution.  Nothing in this kernel may depend on a timer.
template <int NS>
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .u32 n;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra.uni DONE_%=;\n\t"
      "mov.u32 n, 0;\n"
      "SPIN_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra.uni DONE_%=;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 q, n, %3;\n\t"
      "@q bra.uni SPIN_%=;\n\t"
      "trap;\n"
      "DONE_%=:\n\t}"
      ::"r"(addr), "r"(parity), "r"(2000u), "r"(1u << 23)
      : "memory");
}
constexpr int kNsLane = 200, kNsXf = 100, kNsStencil = 0;   // producer / converter lanes, transposers, stencil

struct Params {
  float* partial;        // [N, tiles, Cs] or nullptr
  int N, T, Ho, Wo, Cs;
  int tiles_w, tiles;    // spatial tiles per frame
  int chunks;            // 16-channel chunks
  int pad_h, pad_w;
  int act;               // 1: swish applied to the output
};

// MC = clips per item.  MC = 1: the 32 / LW row strips of a warp are consecutive rows of ONE frame tile.
// MC > 1 (= 32 / LW; frames no larger than LW x Q): every strip is a different CLIP, each with its own
// halo block, so that a warp still has 32 busy lanes and Q rows per thread on 8x8 frames.
template <int S, int LW, int Q, int MC>
struct Geo {
  static constexpr int LS = 32 / LW;                 // row strips per warp
  static_assert(MC == 1 || MC == LS, "one clip per strip");
  static constexpr int TH = MC > 1 ? Q : LS * Q;     // output rows per (clip's) tile
  static constexpr int BW = (LW - 1) * S + 3;        // halo tile
  static constexpr int BH = (TH - 1) * S + 3;
  static constexpr int RB = (Q - 1) * S + 3;         // input rows one thread touches
  static constexpr int raw_clip = BH * BW * kCh * 2; // one TMA box
  static constexpr int raw_bytes = (MC * raw_clip + 127) / 128 * 128;
  // fp32 pairs of one clip's halo block; with MC > 1 the blocks are padded so that neighbouring strips sit
  // 16 banks apart (a half-warp = two strips reads 2 x 64 bytes)
  static constexpr int blk_px = BH * BW;
  static constexpr int blk_bytes = MC > 1 ? (blk_px * 2 + (48 - blk_px * 2 % 32) % 32) * 4 : blk_px * 8;
  static constexpr int plane_bytes = MC * blk_bytes + 16;   // +16: planes 4 apart land on other banks
  static constexpr int ring_bytes = (kPairs * plane_bytes + 127) / 128 * 128;
  static constexpr int stage_clip = TH * LW * kCh * 2;      // dense [TH][LW][16] bf16: one TMA store box
  static constexpr int stage_bytes = MC * stage_clip;
  // planar output frame: [pair][MC][TH*LW] bf16x2 words; +4 words: planes 4 apart sit 16 banks apart, so the
  // converter's two lanes per pixel (pairs j and j+4) never collide; with MC > 1 a clip's block is padded by
  // 8 words so that the four strips of a warp store to different banks
  static constexpr int oblk_bytes = MC > 1 ? (TH * LW + 8) * 4 : TH * LW * 4;
  static constexpr int oplane_bytes = MC * oblk_bytes + 16;
  static constexpr int pout_bytes = (kPairs * oplane_bytes + 127) / 128 * 128;
  static constexpr int off_raw = 256;
  static constexpr int off_ring = off_raw + kRaw * raw_bytes;
  static constexpr int off_pout = off_ring + kRing * ring_bytes;
  static constexpr int off_stage = off_pout + kOut * pout_bytes;
  static constexpr int kStage = MC > 1 ? 1 : 2;             // NDHWC tiles (MC > 1: one, for the shared-memory budget)
  static constexpr int smem = off_stage + kStage * stage_bytes + 128;
  static_assert(smem <= 113 * 1024, "two CTAs per SM");
};

template <int S, int LW, int Q, bool ACT, int MC>
__global__ void __launch_bounds__(kThreads, 2)
dw_planar_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
                 const Params p) {
  using G = Geo<S, LW, Q, MC>;
  extern __shared__ __align__(128) uint8_t dwp_smem_raw[];
  const uint32_t raw_a = smem_u32(dwp_smem_raw);
  const uint32_t smem_s = (raw_a + 127u) & ~127u;
  uint8_t* smem = dwp_smem_raw + (smem_s - raw_a);
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem);   // [kRaw]  TMA landed
  uint64_t* raw_empty = raw_full + kRaw;                    // [kRaw]  transposers done with the tile
  uint64_t* ring_full = raw_empty + kRaw;                   // [kRing] planar frame written
  uint64_t* ring_empty = ring_full + kRing;                 // [kRing] planar frame read by every stencil warp
  uint64_t* out_full = ring_empty + kRing;                  // [kOut]  planar output frame written by every stencil warp
  uint64_t* out_empty = out_full + kOut;                    // [kOut]  converter has read the planar frame
  const uint32_t rawb_s = smem_s + G::off_raw;
  const uint32_t ring_s = smem_s + G::off_ring;
  const uint32_t pout_s = smem_s + G::off_pout;
  const uint32_t stage_s = smem_s + G::off_stage;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int total = ((p.N + MC - 1) / MC) * p.tiles * p.chunks;        // items: (clip group, tile, chunk)

  if (tid == 0) {
    prefetch_tmap(&tmIn);
    prefetch_tmap(&tmOut);
    for (int s = 0; s < kRaw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kXfWarps); }
    for (int s = 0; s < kRing; ++s) { mbar_init(&ring_full[s], kXfWarps); mbar_init(&ring_empty[s], kStWarps); }
    for (int s = 0; s < kOut; ++s) { mbar_init(&out_full[s], kStWarps); mbar_init(&out_empty[s], 1); }
    fence_barrier_init();
  }
  __syncthreads();
  // (launch_pdl) input tiles, output stores and the SE sums all come after this point; every CTA of the
  // persistent grid is resident, so the next kernel may be scheduled as soon as SMs free up
  pdl_wait();
  pdl_trigger();

  // item -> (clip, tile, chunk), the chunk fastest: neighbouring CTAs share the input tile in L2
  auto decode = [&](int idx, int& n, int& tile, int& chunk, int& ho0, int& wo0) {
    chunk = idx % p.chunks;
    const int r = idx / p.chunks;
    n = r / p.tiles;
    tile = r - n * p.tiles;
    n *= MC;                                            // first clip of the item
    const int th = tile / p.tiles_w;
    ho0 = th * G::TH;
    wo0 = (tile - th * p.tiles_w) * LW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one lane)
    reg_dec<kRegsHelper>();
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
        int n, tile, chunk, ho0, wo0;
        decode(idx, n, tile, chunk, ho0, wo0);
        const int hi0 = ho0 * S - p.pad_h, wi0 = wo0 * S - p.pad_w;
        for (int f = 0; f < p.T; ++f) {
          mbar_wait_sleep<kNsLane>(&raw_empty[s], ph ^ 1u);
          // (a clip index past N is out of the tensor: the box arrives as zeros and still counts its bytes)
          mbar_expect_tx(&raw_full[s], static_cast<uint32_t>(MC * G::raw_clip));
#pragma unroll
          for (int c = 0; c < MC; ++c)
            tma_load_5d(rawb_s + s * G::raw_bytes + c * G::raw_clip, &tmIn, chunk * kCh, wi0, hi0, f, n + c, &raw_full[s]);
          if (++s == kRaw) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ output converter + TMA store
    reg_dec<kRegsHelper>();
    // unit = (pixel, half): lane reads the bf16x2 words of 4 channel pairs at one pixel and writes them
    // as 16 contiguous bytes of the NDHWC tile.  Two NDHWC tiles alternate; a tile is rewritten once
    // the bulk store issued from it two frames ago has read it (wait_group.read 1).
    constexpr int kOUnits = MC * G::TH * LW * 2;
    static_assert(kOUnits % 32 == 0, "whole warps of units");
    static_assert(MC == 1 || (G::TH * LW) % 16 == 0, "a warp round of 16 pixels stays inside one clip");
    const uint32_t lsrc = static_cast<uint32_t>((lane & 1) * 4) * G::oplane_bytes + static_cast<uint32_t>(lane >> 1) * 4;
    const uint32_t ldst = static_cast<uint32_t>(lane) * 16;
    int k = 0, j = 0;
    uint32_t kph = 0;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
      int n, tile, chunk, ho0, wo0;
      decode(idx, n, tile, chunk, ho0, wo0);
      for (int f = 0; f < p.T; ++f) {
        if (lane == 0) tma_store_wait_read<G::kStage - 1>();
        mbar_wait_sleep<kNsXf>(&out_full[k], kph);
        __syncwarp();
        const uint32_t src = pout_s + k * G::pout_bytes + lsrc, dst = stage_s + j * G::stage_bytes + ldst;
#pragma unroll 4
        for (int i = 0; i < kOUnits / 32; ++i) {
          // 16 pixels per warp round; with MC > 1 the clips' blocks of the planar frame are padded
          const uint32_t a = src + (MC > 1 ? (i * 16 / (G::TH * LW)) * G::oblk_bytes + (i * 16 % (G::TH * LW)) * 4 : i * 64);
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w0) : "r"(a));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w1) : "r"(a + G::oplane_bytes));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w2) : "r"(a + 2 * G::oplane_bytes));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w3) : "r"(a + 3 * G::oplane_bytes));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + i * 512), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&out_empty[k]);
#pragma unroll
          for (int c = 0; c < MC; ++c)                        // (a clip past N: the whole box is clipped)
            tma_store_5d(&tmOut, stage_s + j * G::stage_bytes + c * G::stage_clip, chunk * kCh, wo0, ho0, f, n + c);
          tma_store_commit();
        }
        if (G::kStage > 1) j ^= 1;
        if (++k == kOut) { k = 0; kph ^= 1u; }
      }
    }
    if (lane == 0) tma_store_wait_read<0>();
  } else if (warp < kStWarp0) {
    // ------------------------------------------------------------ transposers: raw NDHWC tile -> fp32 planar ring
    // unit = (pixel, half): 8 channels = 16 bytes of the raw tile -> 4 channel pairs, each written as
    // one fp32 pair at [pair][pixel]
    reg_dec<kRegsHelper>();
    const int tt = tid - kXfWarp0 * 32;
    constexpr int kUnits = MC * G::blk_px * 2;
    int s = 0, r = 0;
    uint32_t sph = 0, rph = 0;
    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
      for (int f = 0; f < p.T; ++f) {
        mbar_wait_sleep<kNsXf>(&raw_full[s], sph);
        mbar_wait_sleep<kNsXf>(&ring_empty[r], rph ^ 1u);
        const uint32_t src = rawb_s + s * G::raw_bytes, dst = ring_s + r * G::ring_bytes;
#pragma unroll 2
        for (int u = tt; u < kUnits; u += kXfWarps * 32) {
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(src + u * 16));
          const int px = u >> 1;
          const uint32_t d = dst + static_cast<uint32_t>((u & 1) * 4) * G::plane_bytes +
                             (MC > 1 ? static_cast<uint32_t>(px / G::blk_px) * G::blk_bytes + static_cast<uint32_t>(px % G::blk_px) * 8
                                     : static_cast<uint32_t>(px) * 8);
          sts2_f32(d, __uint_as_float(w0 << 16), __uint_as_float(w0 & 0xffff0000u));
          sts2_f32(d + G::plane_bytes, __uint_as_float(w1 << 16), __uint_as_float(w1 & 0xffff0000u));
          sts2_f32(d + 2 * G::plane_bytes, __uint_as_float(w2 << 16), __uint_as_float(w2 & 0xffff0000u));
          sts2_f32(d + 3 * G::plane_bytes, __uint_as_float(w3 << 16), __uint_as_float(w3 & 0xffff0000u));
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&raw_empty[s]);
          mbar_arrive(&ring_full[r]);
        }
        if (++s == kRaw) { s = 0; sph ^= 1u; }
        if (++r == kRing) { r = 0; rph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------ stencil warps: warp = channel pair
    // The pair index goes through a shuffle so that ptxas can prove it warp-uniform: the constant-bank
    // loads below then land in uniform registers and FFMA2 takes them as its uniform operand.
    reg_inc<kRegsStencil>();
    const int pl = __shfl_sync(0xffffffffu, warp - kStWarp0, 0);       // plane (pair inside the chunk)
    const int col = lane % LW, strip = lane / LW;
    const uint32_t toff = static_cast<uint32_t>(pl) * G::plane_bytes +
                          (MC > 1 ? static_cast<uint32_t>(strip) * G::blk_bytes + static_cast<uint32_t>(col * S) * 8
                                  : static_cast<uint32_t>((strip * Q * S) * G::BW + col * S) * 8);
    const uint32_t soff = static_cast<uint32_t>(pl) * G::oplane_bytes +
                          (MC > 1 ? static_cast<uint32_t>(strip) * G::oblk_bytes + static_cast<uint32_t>(col) * 4
                                  : static_cast<uint32_t>((strip * Q) * LW + col) * 4);
    int r = 0, k = 0;
    uint32_t rph = 0, kph = 0;

    for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
      int n, tile, chunk, ho0, wo0;
      decode(idx, n, tile, chunk, ho0, wo0);
      const int pair = chunk * kPairs + pl;                  // warp-uniform
      const bool chan = 2 * pair < p.Cs;
      const float2* tp = c_taps + (chan ? pair : 0) * 28;
      float2 wv[27];
#pragma unroll
      for (int i = 0; i < 27; ++i) wv[i] = tp[i];
      const float2 bia = tp[27];
      const bool col_ok = wo0 + col < p.Wo;
      // output rows of this thread inside the image (MC > 1: the strip is its own clip, which may lie past N)
      const int rows_ok = MC > 1 ? (n + strip < p.N ? p.Ho - ho0 : 0) : p.Ho - (ho0 + strip * Q);

      // The BN shift is added when a frame is emitted (FADD2 with the uniform operand), so a set is
      // (re)started by its first tap with a zero addend and the shift never has to sit in registers.
      float2 acc[3][Q];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[a][q] = make_float2(0.f, 0.f);
      float2 ssum = make_float2(0.f, 0.f);

      auto emit_out = [&](float2 (&A)[Q]) {
        mbar_wait_sleep<kNsStencil>(&out_empty[k], kph ^ 1u);
        const uint32_t dst = pout_s + k * G::pout_bytes + soff;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          float2 v = __fadd2_rn(A[q], bia);
          if (ACT) {
            // swish(a) = h + h tanh(h), h = a / 2
            const float2 h = __fmul2_rn(v, make_float2(0.5f, 0.5f));
            float2 t;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
            v = __ffma2_rn(h, t, h);
          }
          const __nv_bfloat162 hb = __float22bfloat162_rn(v);
          sts_b32(dst + q * (LW * 4), *reinterpret_cast<const uint32_t*>(&hb));
          if (col_ok && q < rows_ok) ssum = __fadd2_rn(ssum, v);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_full[k]);
        if (++k == kOut) { k = 0; kph ^= 1u; }
      };

      // One step: input frame t contributes tap dt=0 to output t+1 (set A0, restarted here), dt=1 to
      // output t (A1) and dt=2 to output t-1 (A2), which is complete afterwards.
      auto step = [&](bool first, float2 (&A0)[Q], float2 (&A1)[Q], float2 (&A2)[Q]) {
        mbar_wait_sleep<kNsStencil>(&ring_full[r], rph);
        const uint32_t base = ring_s + r * G::ring_bytes + toff;
#pragma unroll
        for (int rr = 0; rr < G::RB; ++rr) {
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) {
            const float2 x = lds2_f32(base + (rr * G::BW + dw) * 8);
#pragma unroll
            for (int dh = 0; dh < 3; ++dh) {
              const int qn = rr - dh;                        // = q * S for the output row q it feeds
              if (qn >= 0 && qn % S == 0 && qn / S < Q) {    // resolved at compile time
                const int q = qn / S;
                A0[q] = fma2(x, wv[(0 * 3 + dh) * 3 + dw], (dh == 0 && dw == 0) ? make_float2(0.f, 0.f) : A0[q]);
                A1[q] = fma2(x, wv[(1 * 3 + dh) * 3 + dw], A1[q]);
                A2[q] = fma2(x, wv[(2 * 3 + dh) * 3 + dw], A2[q]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring_empty[r]);
        if (++r == kRing) { r = 0; rph ^= 1u; }
        if (!first) emit_out(A2);
      };

      int t = 0;
      for (; t + 3 <= p.T; t += 3) {
        step(t == 0, acc[1], acc[0], acc[2]);
        step(false, acc[2], acc[1], acc[0]);
        step(false, acc[0], acc[2], acc[1]);
      }
      const int rem = p.T - t;
      if (rem >= 1) step(t == 0, acc[1], acc[0], acc[2]);
      if (rem == 2) step(false, acc[2], acc[1], acc[0]);
      if (rem == 0) emit_out(acc[2]);
      else if (rem == 1) emit_out(acc[0]);
      else emit_out(acc[1]);

      if (p.partial != nullptr) {
        // MC = 1: the whole warp is one clip's tile; MC > 1: every LW lanes are a clip of their own
#pragma unroll
        for (int o = (MC > 1 ? LW / 2 : 16); o > 0; o >>= 1) {
          ssum.x += __shfl_xor_sync(0xffffffffu, ssum.x, o);
          ssum.y += __shfl_xor_sync(0xffffffffu, ssum.y, o);
        }
        const int nn = MC > 1 ? n + strip : n;
        if ((MC > 1 ? col == 0 : lane == 0) && chan && nn < p.N) {
          float* dstp = p.partial + (static_cast<long>(nn) * p.tiles + tile) * p.Cs + 2 * pair;
          dstp[0] = ssum.x;
          dstp[1] = ssum.y;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------ host
struct Plan { int LW, Q, TH, MC, tiles_w, tiles_h, chunks; size_t smem; };

static Plan make_plan(int H, int W, int Cs, int stride) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  Plan pl{};
  // lanes along the width: 32 columns where the image has them, else 16 / 8 with 2 / 4 row strips per warp
  pl.LW = Wo > 16 ? 32 : (Wo > 8 ? 16 : 8);
  pl.Q = pl.LW == 32 ? 8 : (pl.LW == 16 ? 8 : 2);
  if (pl.LW == 16 && Ho <= 8) pl.Q = 4;
  if (stride == 1 && pl.LW >= 16 && Ho > 8) {
    // 8 or 7 output rows per thread, whichever pads the height less (56, 28, 14 rows: 7)
    const int th8 = (32 / pl.LW) * 8, th7 = (32 / pl.LW) * 7;
    const int pad8 = (Ho + th8 - 1) / th8 * th8, pad7 = (Ho + th7 - 1) / th7 * th7;
    pl.Q = pad7 < pad8 ? 7 : 8;
  }
  // stride 2: the halo tile (and with it the fp32 planar ring) is four times the output tile; two output
  // rows per thread keep two CTAs resident per SM
  if (stride == 2) pl.Q = 2;
  pl.MC = 1;
  pl.TH = (32 / pl.LW) * pl.Q;
  // stride 1 on frames of at most 8x8 pixels: the four strips of a warp are four CLIPS, 8 rows per thread
  if (stride == 1 && pl.LW == 8 && Ho <= 8) { pl.MC = 4; pl.Q = 8; pl.TH = 8; }
  pl.tiles_w = (Wo + pl.LW - 1) / pl.LW;
  pl.tiles_h = (Ho + pl.TH - 1) / pl.TH;
  pl.chunks = (Cs + kCh - 1) / kCh;
  return pl;
}

template <int S, int LW, int Q, bool ACT, int MC = 1>
static int launch(const CUtensorMap& ti, const CUtensorMap& to, const Params& p, int N, const Plan& pl, cudaStream_t st) {
  using G = Geo<S, LW, Q, MC>;
  auto kern = dw_planar_kernel<S, LW, Q, ACT, MC>;
  static SmemOptIn optin;
  const cudaError_t e = ensure_dynamic_smem(kern, optin, G::smem);
  if (e != cudaSuccess) {
    set_error("x3d_dw3x3x3_planar_fwd: smem attribute (%d B): %s", G::smem, cudaGetErrorString(e));
    return X3D_ERR_LAUNCH;
  }
  const long total = (long)((N + MC - 1) / MC) * pl.tiles_w * pl.tiles_h * pl.chunks;
  long gx = 2L * device_sm_count();
  if (gx > total) gx = total;
  const cudaError_t le = launch_pdl(kern, dim3((unsigned)gx), dim3(kThreads), G::smem, st, ti, to, p);
  if (le != cudaSuccess) {
    set_error("x3d_dw3x3x3_planar_fwd: launch: %s", cudaGetErrorString(le));
    return X3D_ERR_LAUNCH;
  }
  return check_launch("x3d_dw3x3x3_planar_fwd");
}

}  // namespace dwp
}  // namespace x3d

using namespace x3d;

extern "C" int x3d_dw_planar_partial_blocks(int T, int H, int W, int C, int stride) {
  if (T <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || (stride != 1 && stride != 2)) return 0;
  const dwp::Plan pl = dwp::make_plan(H, W, C, stride);
  return pl.tiles_w * pl.tiles_h;
}

extern "C" int x3d_dw_planar_lane_permille(int T, int H, int W, int C, int stride) {
  if (T <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || (stride != 1 && stride != 2)) return 0;
  const dwp::Plan pl = dwp::make_plan(H, W, C, stride);
  const long Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  return (int)(1000 * Ho * Wo / ((long)pl.tiles_w * pl.LW * pl.tiles_h * pl.TH));
}

// taps: device pointer to [ceil(C/2)][28] float2 = per channel pair the 27 BN-folded taps (dt, dh, dw
// major) and the BN shift, as (channel 2p, channel 2p+1)
extern "C" int x3d_dw3x3x3_planar_fwd(const void* in, const float* taps, void* out, float* se_partial, int N, int T,
                                      int H, int W, int C, int stride, int pad_h, int pad_w, int act, void* stream) {
  X3D_REQUIRE(in && taps && out, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_planar_fwd: null pointer");
  X3D_REQUIRE(act == 0 || (act == 1 && se_partial == nullptr), X3D_ERR_INVALID_ARG,
              "x3d_dw3x3x3_planar_fwd: act=%d (0, or 1 = swish without SE sums)", act);
  X3D_REQUIRE(C > 0 && C % 8 == 0 && C / 2 <= dwp::kMaxPairs, X3D_ERR_UNSUPPORTED,
              "x3d_dw3x3x3_planar_fwd: C=%d (multiple of 8, at most %d)", C, 2 * dwp::kMaxPairs);
  X3D_REQUIRE(stride == 1 || stride == 2, X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_planar_fwd: stride %d", stride);
  X3D_REQUIRE(N > 0 && T > 0 && H > 0 && W > 0, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_planar_fwd: bad extent");
  X3D_REQUIRE(pad_h >= 0 && pad_h <= 1 && pad_w >= 0 && pad_w <= 1, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_planar_fwd: pad_before must be 0 or 1");
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(taps) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_planar_fwd: pointers must be 16-byte aligned");
  X3D_REQUIRE(device_sm_count() > 0 && device_is_sm100(), X3D_ERR_NO_DEVICE, "x3d_dw3x3x3_planar_fwd: needs an sm_100 device");
  EncodeTiledFn enc = tensor_map_encoder();
  X3D_REQUIRE(enc != nullptr, X3D_ERR_NO_DEVICE, "x3d_dw3x3x3_planar_fwd: cuTensorMapEncodeTiled unavailable");
  const dwp::Plan pl = dwp::make_plan(H, W, C, stride);
  X3D_REQUIRE((long)N * pl.tiles_w * pl.tiles_h * pl.chunks < (1L << 31), X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_planar_fwd: too many work items");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // the layer's tap table into the constant bank (stream-ordered: after the previous launch that reads
  // the table, before this one)
  void* sym = nullptr;
  cudaError_t ce = cudaGetSymbolAddress(&sym, dwp::c_taps);
  X3D_REQUIRE(ce == cudaSuccess, X3D_ERR_LAUNCH, "x3d_dw3x3x3_planar_fwd: constant table: %s", cudaGetErrorString(ce));
  ce = cudaMemcpyAsync(sym, taps, (size_t)(C / 2) * 28 * sizeof(float2), cudaMemcpyDeviceToDevice, st);
  X3D_REQUIRE(ce == cudaSuccess, X3D_ERR_LAUNCH, "x3d_dw3x3x3_planar_fwd: tap upload: %s", cudaGetErrorString(ce));

  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const int BW = (pl.LW - 1) * stride + 3, BH = (pl.TH - 1) * stride + 3;
  CUtensorMap ti, to;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2, (cuuint64_t)T * H * W * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)dwp::kCh, (cuuint32_t)BW, (cuuint32_t)BH, 1, 1};
    CUresult r = enc(&ti, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_dw3x3x3_planar_fwd: input tensor map failed (%d) box [16,%d,%d]", (int)r, BW, BH);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)Wo * C * 2, (cuuint64_t)Ho * Wo * C * 2, (cuuint64_t)T * Ho * Wo * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)dwp::kCh, (cuuint32_t)pl.LW, (cuuint32_t)pl.TH, 1, 1};
    CUresult r = enc(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_dw3x3x3_planar_fwd: output tensor map failed (%d)", (int)r);
  }
  dwp::Params p;
  p.partial = se_partial; p.N = N; p.T = T; p.Ho = Ho; p.Wo = Wo; p.Cs = C;
  p.tiles_w = pl.tiles_w; p.tiles = pl.tiles_w * pl.tiles_h; p.chunks = pl.chunks;
  p.pad_h = pad_h; p.pad_w = pad_w; p.act = act;
#define X3D_DWP(SS, LWW, QQ) \
  if (stride == SS && pl.LW == LWW && pl.Q == QQ) \
    return act ? dwp::launch<SS, LWW, QQ, true>(ti, to, p, N, pl, st) : dwp::launch<SS, LWW, QQ, false>(ti, to, p, N, pl, st)
  if (pl.MC == 4)
    return act ? dwp::launch<1, 8, 8, true, 4>(ti, to, p, N, pl, st) : dwp::launch<1, 8, 8, false, 4>(ti, to, p, N, pl, st);
  X3D_DWP(1, 32, 8); X3D_DWP(1, 16, 8); X3D_DWP(1, 32, 7); X3D_DWP(1, 16, 7); X3D_DWP(1, 16, 4); X3D_DWP(1, 8, 2);
  X3D_DWP(2, 32, 2); X3D_DWP(2, 16, 2); X3D_DWP(2, 8, 2);
#undef X3D_DWP
  set_error("x3d_dw3x3x3_planar_fwd: no kernel for stride=%d LW=%d Q=%d", stride, pl.LW, pl.Q);
  return X3D_ERR_UNSUPPORTED;
}
