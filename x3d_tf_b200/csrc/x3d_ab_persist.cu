// Fused bottleneck front half, persistent / warp-specialised form:
//   1x1x1 expand conv + BN + ReLU  ->  channelwise 3x3x3 conv + BN (+ swish) (+ SE partial sums)
// = Bottleneck.a / bn_a / relu / b / bn_b (+ the reduction of se_pool, + the swish of blocks
// without SE), reference model.py:306-316, as ONE kernel.  The `inner`-wide tensor between the two
// convolutions never leaves the SM, and -- unlike the unfused pair -- it is never rounded to bf16:
// the stencil reads the fp32 accumulator values.
//
// One CTA per SM (640 threads = 5 warpgroups), fixed to one chunk of CH inner channels (blockIdx.y), walks a
// list of work items (clip x spatial tile of Q x SW outputs).  Per item it marches over the T
// frames; the five roles run decoupled through mbarrier rings, across item boundaries too:
//   warp 0  lane 0   TMA producer: halo tile of the block INPUT for (item, frame) into an XS-deep
//                    ring (BH x BW pixels x 64-channel chunks, 128B swizzle; pixels and channels
//                    outside the tensor are zero-filled by the TMA unit);
//   warp 1  lane 0   tcgen05.mma issuer: per frame one ones x shift MMA (the bn_a shift enters the
//                    accumulator through the tensor core) + ceil(Cin/16) K=16 MMAs per 128-pixel
//                    block against the CTA's resident weight slice; result in one of two TMEM
//                    accumulators;
//   warps 4-7        drain: TMEM -> registers (tcgen05.ld, one lane quarter per warp) -> ReLU,
//                    pixels outside the image forced to zero (TF 'SAME' pads the OUTPUT of `a`)
//                    -> fp32 frame ring in shared memory (3 slots);
//   warps 8-19       stencil: one channel pair x one output row per thread; every ring value is
//                    read once per frame (LDS.64, no unpacking) and scattered into three rotating
//                    accumulator sets (output frames t-1, t, t+1) with packed FFMA2; the finished
//                    frame is rounded to bf16 into one of two staging buffers;
//   warp 2  lane 0   TMA store of staged output frames (clipped at the tensor edge).
// The stencil warps do nothing but LDS + FFMA2 + one barrier wait/arrive pair per frame: the
// channelwise conv is bound by the fp32 FMA pipe (27 MACs per output against 2 bytes of HBM
// traffic once the expand tensor stays on chip), so their instruction stream is what sets the
// kernel's speed.  Registers follow the roles (setmaxnreg): the kernel launches with 96 per
// thread; the utility warpgroup drops to 40, the drain warpgroup to 56, and the three stencil
// warpgroups (three warps per scheduler) rise to 128.  Every barrier wait carries a suspend-time
// hint so that a waiting role sleeps in hardware instead of taking issue slots from the stencil.
#include <stdlib.h>

#include "tma_common.cuh"

namespace x3d {
namespace abp {

using namespace ptx;

constexpr int kThreads = 640;
constexpr int kDrainWarp0 = 4, kDrainWarps = 4;
constexpr int kStencilWarp0 = 8, kStencilWarps = 12, kStencilThreads = kStencilWarps * 32;
constexpr int kMaxQ = 13;         // output rows per tile (= stencil threads / channel pairs of the narrowest chunk)
constexpr int kRegsUtil = 40, kRegsDrain = 56, kRegsStencil = 128;   // 128*40 + 128*56 + 384*128 = 640*96
constexpr int kFrames = 3;        // fp32 frame ring depth
constexpr int kOut = 2;           // output staging buffers
constexpr int kMaxAcc = 4;        // TMEM accumulator buffers (MB*CHN columns each, 512 columns in all)
constexpr int kMaxXS = 4;

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
template <int N> __device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
// Barrier wait that sleeps in hardware: every try_wait may stay suspended up to ~4 us, the thread
// wakes when the phase completes.  A lost arrival still traps (after seconds) instead of hanging.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns) {
  if (ns == 0) { mbar_wait_lean(bar, parity); return; }       // plain polling (experiments)
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
      ::"r"(addr), "r"(parity), "r"(ns), "r"(1u << 24)
      : "memory");
}
// Arrive without release semantics: for hand-overs whose only obligation is "my READS of the
// buffer are done" (they are: the loaded values were consumed by arithmetic before this point).
// A releasing arrive would also wait for the thread's outstanding global stores to be performed.
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
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
__device__ __forceinline__ float2 lds2_f32(uint32_t a) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts2_f32(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void sts2_bf16(uint32_t a, float2 v) {
  __nv_bfloat162 h = __float22bfloat162_rn(v);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(*reinterpret_cast<uint32_t*>(&h)) : "memory");
}
__device__ __forceinline__ void sts4_relu(uint32_t a, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(fmaxf(__uint_as_float(v0), 0.f)),
               "f"(fmaxf(__uint_as_float(v1), 0.f)), "f"(fmaxf(__uint_as_float(v2), 0.f)),
               "f"(fmaxf(__uint_as_float(v3), 0.f))
               : "memory");
}
__device__ __forceinline__ void sts4_zero(uint32_t a) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
}

struct Params {
  const float* bias_a;   // [Cs] BN shift of bn_a
  const float* w;        // [27, Cs] BN-folded channelwise taps
  const float* bias;     // [Cs] BN shift of bn_b
  float* partial;        // [N, tiles, Cs] or nullptr
  int N, T, H, W, Ho, Wo, Cs;
  int Q, tiles_w, tiles;
  int pad_h, pad_w;
  int KC;                // 64-wide K chunks of the expand GEMM
  int k16;               // K=16 MMA steps in total (= ceil(Cin/16))
  int MB;                // 128-pixel blocks of the halo tile
  int CHN;               // MMA N (CH rounded up to 16)
  int XS;                // depth of the x-tile ring
  int NA, acc_cols;      // TMEM accumulator buffers and their column pitch
  int x_kc_bytes;        // one K chunk of an x stage (MB * 16 KiB: whole 128-row blocks)
  int x_stage_bytes;     // KC * x_kc_bytes
  int x_box_bytes;       // bytes one x-tile TMA box delivers (per K chunk)
  int slot_bytes;        // frame ring slot
  int stage_bytes;       // output staging buffer
  int off_wa, off_x, off_ring, off_stage, off_red, off_w2, off_bias;   // from the 1024-aligned base
  int act;               // 1: swish applied to the output (blocks without SE, model.py:316)
  uint32_t sleep_ns;     // suspend-time hint of the barrier waits
  int dbg;               // X3D_ABP_DEBUG timing experiments (0 in production; results are garbage otherwise):
                         //   1 = stencil warps alone (no barrier waits, other roles idle), 2 = ... without staging
};

// Pixel pitch of the fp32 frame ring: CH floats, padded so that pitch/16 is odd (32 lanes = 32
// consecutive pixels then write their 16-byte vectors conflict-free).
template <int CH> struct RingPitch { static constexpr int value = ((CH * 4 / 16) & 1) ? CH * 4 : CH * 4 + 16; };

template <int S, int SW, int CH>
__global__ void __launch_bounds__(kThreads, 1)
ab_persist_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmOut, const Params p) {
  constexpr int BW = (SW - 1) * S + 3;
  constexpr int PS = RingPitch<CH>::value;   // bytes per ring pixel
  constexpr int RS = BW * PS;                // bytes per ring row
  constexpr int OPS = CH * 2;                // bytes per staged output pixel (dense: TMA store)
  constexpr int C2 = CH / 2;
  constexpr int kRegTaps = 18;               // dt = 0, 1 taps in registers; dt = 2 in shared memory

  extern __shared__ __align__(1024) uint8_t abp_smem_raw[];
  const uint32_t raw_s = smem_u32(abp_smem_raw);
  const uint32_t smem_s = (raw_s + 1023u) & ~1023u;
  uint8_t* smem = abp_smem_raw + (smem_s - raw_s);
  // barriers live in the first 512 bytes
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem);     // weights landed
  uint64_t* x_full = w_full + 1;                            // [XS] x tile landed
  uint64_t* x_empty = x_full + kMaxXS;                      // [XS] MMAs reading the stage retired
  uint64_t* acc_full = x_empty + kMaxXS;                    // [NA] TMEM accumulator complete
  uint64_t* acc_empty = acc_full + kMaxAcc;                 // [NA] TMEM accumulator drained
  uint64_t* ring_full = acc_empty + kMaxAcc;                // [3] frame written by every drain warp
  uint64_t* ring_empty = ring_full + kFrames;               // [3] frame read by every stencil warp
  uint64_t* out_full = ring_empty + kFrames;                // [2] output frame packed by every stencil warp
  uint64_t* out_empty = out_full + kOut;                    // [2] TMA store has read the buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_empty + kOut);
  const uint32_t wa_s = smem_s + p.off_wa;
  const uint32_t x_s = smem_s + p.off_x;
  const uint32_t ring_s = smem_s + p.off_ring;
  const uint32_t stage_s = smem_s + p.off_stage;
  float* s_red = reinterpret_cast<float*>(smem + p.off_red);
  const uint32_t w2_s = smem_s + p.off_w2;
  // bias as a GEMM operand: ones[128 x 16] (one 8-row atom, SBO = 0) x  [hi(shift), lo(shift)] rows
  const uint32_t ones_s = smem_s + p.off_bias;               // 1 KiB atom
  const uint32_t biasop_s = ones_s + 1024;                   // [CHN rows x 128 B], 128B swizzle

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.y * CH;
  const int total = p.N * p.tiles;                           // work items of this channel chunk
  const int BH = (p.Q - 1) * S + 3;
  const int npix = BH * BW;

  // The BN shift of bn_a enters the accumulator through the tensor core: D = ones x Bop with
  // ones[m, 0:2] = 1 and Bop[n, 0:2] = (hi, lo) bf16 split of shift[n] (exact to 2^-17).  Both
  // tiles are K-major, 128B-swizzled: the 16-byte chunk j of row r sits at chunk position j ^ (r % 8).
  for (int i = tid; i < 8 + p.CHN; i += kThreads) {
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
    for (int s = 0; s < kMaxXS; ++s) {
      mbar_init(&x_full[s], 1);
      mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], kDrainWarps);
    }
    for (int s = 0; s < kFrames; ++s) {
      mbar_init(&ring_full[s], kDrainWarps);
      mbar_init(&ring_empty[s], kStencilWarps);
    }
    for (int s = 0; s < kOut; ++s) {
      mbar_init(&out_full[s], kStencilWarps);
      mbar_init(&out_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const bool solo = (p.dbg & 3) != 0;                               // timing experiment: stencil warps alone
  if (warp < kDrainWarp0) {
    reg_dec<kRegsUtil>();
  if (solo) {
  } else if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      const int w_chunk_bytes = p.CHN * 128;
      mbar_expect_tx(w_full, static_cast<uint32_t>(p.KC * w_chunk_bytes));
      for (int kc = 0; kc < p.KC; ++kc)
        tma_load_2d_s(wa_s + kc * w_chunk_bytes, &tmW, kc * 64, c0, w_full);
      int s = 0;
      uint32_t ph = 0;
      for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
        const int n = idx / p.tiles, tile = idx - n * p.tiles;
        const int tile_h = tile / p.tiles_w, tile_w = tile - tile_h * p.tiles_w;
        const int hi0 = tile_h * p.Q * S - p.pad_h, wi0 = tile_w * SW * S - p.pad_w;
        for (int f = 0; f < p.T; ++f) {
          mbar_wait_sleep(&x_empty[s], ph ^ 1u, p.sleep_ns);
          mbar_expect_tx(&x_full[s], static_cast<uint32_t>(p.KC * p.x_box_bytes));
          for (int kc = 0; kc < p.KC; ++kc)
            tma_load_5d(x_s + s * p.x_stage_bytes + kc * p.x_kc_bytes, &tmX, kc * 64, wi0, hi0, f, n, &x_full[s]);
          if (++s == p.XS) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      const int w_chunk_bytes = p.CHN * 128;
      const uint32_t idesc = make_idesc_bf16(p.CHN);
      mbar_wait_sleep(w_full, 0, p.sleep_ns);
      int s = 0, b = 0;
      uint32_t ph = 0, bph = 0;
      for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
        for (int f = 0; f < p.T; ++f) {
          mbar_wait_sleep(&x_full[s], ph, p.sleep_ns);
          mbar_wait_sleep(&acc_empty[b], bph ^ 1u, p.sleep_ns);
          tcgen05_after_sync();
          const uint32_t x_base = x_s + s * p.x_stage_bytes;
          for (int mb = 0; mb < p.MB; ++mb) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(b * p.acc_cols + mb * p.CHN);
            umma_bf16(d_tmem, make_desc_sw128(ones_s, 0), make_desc_sw128(biasop_s), idesc, 0u);
            for (int k = 0; k < p.k16; ++k) {
              const int kc = k >> 2, kk = k & 3;
              umma_bf16(d_tmem, make_desc_sw128(x_base + kc * p.x_kc_bytes + mb * 16384 + kk * 32),
                        make_desc_sw128(wa_s + kc * w_chunk_bytes + kk * 32), idesc, 1u);
            }
          }
          umma_commit(&x_empty[s]);            // the x stage is free once these MMAs have retired
          umma_commit(&acc_full[b]);
          if (++s == p.XS) { s = 0; ph ^= 1u; }
          if (++b == p.NA) { b = 0; bph ^= 1u; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------ TMA store of staged frames (one lane)
    if (lane == 0) {
      int k = 0;
      uint32_t kph = 0;
      for (int idx = blockIdx.x; idx < total; idx += gridDim.x) {
        const int n = idx / p.tiles, tile = idx - n * p.tiles;
        const int tile_h = tile / p.tiles_w, tile_w = tile - tile_h * p.tiles_w;
        const int ho0 = tile_h * p.Q, wo0 = tile_w * SW;
        for (int f = 0; f < p.T; ++f) {
          mbar_wait_sleep(&out_full[k], kph, p.sleep_ns);
          tma_store_5d(&tmOut, stage_s + k * p.stage_bytes, c0, wo0, ho0, f, n);
          tma_store_commit();
          tma_store_wait_read<0>();            // this thread has nothing else to do: release the buffer at once
          mbar_arrive(&out_empty[k]);
          if (++k == kOut) { k = 0; kph ^= 1u; }
        }
      }
    }
  }
  } else if (warp < kStencilWarp0) {
    // ------------------------------------------------------------ drain warps: TMEM -> fp32 ring
    reg_dec<kRegsDrain>();
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    uint32_t rph = 0, bph = 0;
    int r = 0, b = 0;
    const uint32_t drain_off = static_cast<uint32_t>(q * 32 + lane) * PS;
    const uint32_t drain_tm = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int idx = blockIdx.x; idx < total && !solo; idx += gridDim.x) {
      const int n = idx / p.tiles, tile = idx - n * p.tiles;
      const int tile_h = tile / p.tiles_w, tile_w = tile - tile_h * p.tiles_w;
      const int hi0 = tile_h * p.Q * S - p.pad_h, wi0 = tile_w * SW * S - p.pad_w;
      // per-lane pixel bookkeeping, once per item: for pixel block mb this lane owns halo pixel
      // rr = mb*128 + q*32 + lane; bit mb of pix_m = it exists, of in_m = it is inside the image
      uint32_t pix_m = 0, in_m = 0;
      int mb_n = 0;
      for (int mb = 0; mb < p.MB; ++mb) {
        if (mb * 128 + q * 32 >= npix) break;
        mb_n = mb + 1;
        const int rr = mb * 128 + q * 32 + lane;
        const int hh = rr / BW, ww = rr - hh * BW;
        if (rr < npix) {
          pix_m |= 1u << mb;
          if (static_cast<unsigned>(hi0 + hh) < static_cast<unsigned>(p.H) &&
              static_cast<unsigned>(wi0 + ww) < static_cast<unsigned>(p.W))
            in_m |= 1u << mb;
        }
      }
      for (int f = 0; f < p.T; ++f) {
        mbar_wait_sleep(&acc_full[b], bph, p.sleep_ns);
        tcgen05_after_sync();
        mbar_wait_sleep(&ring_empty[r], rph ^ 1u, p.sleep_ns);
        const uint32_t slot_base = ring_s + r * p.slot_bytes + drain_off;
        const uint32_t tbase = drain_tm + static_cast<uint32_t>(b * p.acc_cols);
        for (int mb = 0; mb < mb_n; ++mb) {
          const bool pix = (pix_m >> mb) & 1u, inside = (in_m >> mb) & 1u;
          const uint32_t taddr = tbase + static_cast<uint32_t>(mb * p.CHN);
          const uint32_t dst = slot_base + static_cast<uint32_t>(mb * 128) * PS;
          // 16 accumulator columns at a time (the drain runs on 56 registers); a software-pipelined
          // form with the next load in flight during the stores measured slower (0.73 against 0.64 ms)
#pragma unroll
          for (int cb = 0; cb < CH; cb += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + cb, v);                              // CHN >= cb + 16: CHN is CH rounded up to 16
            tmem_ld_wait();
            if (inside) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (cb + 4 * g < CH)
                  sts4_relu(dst + (cb + 4 * g) * 4, v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            } else if (pix) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (cb + 4 * g < CH) sts4_zero(dst + (cb + 4 * g) * 4);
            }
          }
        }
        tcgen05_before_sync();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc_empty[b]);
          mbar_arrive(&ring_full[r]);
        }
        if (++r == kFrames) { r = 0; rph ^= 1u; }
        if (++b == p.NA) { b = 0; bph ^= 1u; }
      }
    }
  } else {
    // ------------------------------------------------------------ stencil warps
    reg_inc<kRegsStencil>();
    const int st = tid - kStencilWarp0 * 32;
    const int slot = st / C2, cp = st - slot * C2;
    const int c = c0 + 2 * cp;
    const bool in_slot = slot < p.Q;
    const bool chan = c < p.Cs;
    float2 wr[kRegTaps];
    float2 bia = make_float2(0.f, 0.f);
    if (in_slot && chan) {
#pragma unroll
      for (int i = 0; i < kRegTaps; ++i) wr[i] = ld2(p.w + i * p.Cs + c);
      bia = ld2(p.bias + c);
    } else {
#pragma unroll
      for (int i = 0; i < kRegTaps; ++i) wr[i] = make_float2(0.f, 0.f);
    }
    const uint32_t w2_t = w2_s + static_cast<uint32_t>(cp) * 8;
    if (slot == 0) {
#pragma unroll
      for (int i = 0; i < 27 - kRegTaps; ++i) {
        const float2 w = chan ? ld2(p.w + (kRegTaps + i) * p.Cs + c) : make_float2(0.f, 0.f);
        sts2_f32(w2_t + i * (C2 * 8), w);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kStencilThreads) : "memory");

    const int aslot = in_slot ? slot : 0;       // spare threads stay inside the ring
    const uint32_t toff = static_cast<uint32_t>(aslot * S) * RS + static_cast<uint32_t>(cp) * 8;
    const uint32_t soff = static_cast<uint32_t>(aslot * SW) * OPS + static_cast<uint32_t>(cp) * 4;
    const bool lane0 = lane == 0;
    uint32_t rph = 0, item_par = 0, kph = 0;
    int r = 0, k = 0;

    for (int idx = blockIdx.x; idx < total; idx += gridDim.x, item_par ^= 1u) {
      const int n = idx / p.tiles, tile = idx - n * p.tiles;
      const int tile_h = tile / p.tiles_w, tile_w = tile - tile_h * p.tiles_w;
      const int ho0 = tile_h * p.Q, wo0 = tile_w * SW;
      const bool on = in_slot && chan && ho0 + slot < p.Ho;
      const bool on_se = on && p.partial != nullptr;
      int ncol = p.Wo - wo0;
      if (ncol > SW) ncol = SW;

      // Three accumulator sets rotate over output frames.  A set is (re)started by the first tap of
      // the dt=0 pass with the BN shift as addend; only frame 0 needs a preset.
      float2 acc[3][SW];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int j = 0; j < SW; ++j) acc[a][j] = bia;
      float2 ssum = make_float2(0.f, 0.f);
      // Finished output frame -> bf16 staging buffer (dense [Q][SW][CH]); the store lane sends it
      // with one TMA box store, clipped at the tensor edge.  Shared-memory stores + bulk copy keep
      // the LSU free for the ring reads (direct 4-byte global stores from the accumulators were
      // measured: 0.63 ms against 0.54 ms for the stage-2 layer).
      auto emit_out = [&](float2 (&A)[SW]) {
        if ((p.dbg & 3) == 2) return;
        if (!solo) mbar_wait_sleep(&out_empty[k], kph ^ 1u, p.sleep_ns);
        const uint32_t dst = stage_s + k * p.stage_bytes + soff;
        if (p.act) {
          // swish(a) = h + h tanh(h), h = a / 2: one multiply, two MUFU, one FMA per channel pair
#pragma unroll
          for (int j = 0; j < SW; ++j) {
            const float2 h = __fmul2_rn(A[j], make_float2(0.5f, 0.5f));
            float2 t;
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
            asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
            A[j] = __ffma2_rn(h, t, h);
          }
        }
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
        if (lane0 && !solo) mbar_arrive(&out_full[k]);
        if (++k == kOut) { k = 0; kph ^= 1u; }
      };

      // One step: input frame t contributes tap dt=0 to output t+1 (set A0, restarted here), dt=1 to
      // output t (A1) and dt=2 to output t-1 (A2), which is complete afterwards.
      auto step = [&](bool first, float2 (&A0)[SW], float2 (&A1)[SW], float2 (&A2)[SW]) {
        if (!solo) mbar_wait_sleep(&ring_full[r], rph, p.sleep_ns);
        const uint32_t base = ring_s + r * p.slot_bytes + toff;
#pragma unroll
        for (int dh = 0; dh < 3; ++dh) {
          float2 w2[3];
#pragma unroll
          for (int dw = 0; dw < 3; ++dw) w2[dw] = lds2_f32(w2_t + (dh * 3 + dw) * (C2 * 8));
#pragma unroll
          for (int jj = 0; jj < BW; ++jj) {
            // one ring value feeds up to 3 output columns x 3 output frames, then dies
            const float2 x = lds2_f32(base + dh * RS + jj * PS);
#pragma unroll
            for (int dw = 0; dw < 3; ++dw) {
              const int jn = jj - dw;                       // = j * S for the output column j it feeds
              if (jn >= 0 && jn % S == 0 && jn / S < SW) {  // resolved at compile time
                const int j = jn / S;
                A0[j] = fma2(x, wr[(0 * 3 + dh) * 3 + dw], (dh == 0 && dw == 0) ? bia : A0[j]);
                A1[j] = fma2(x, wr[(1 * 3 + dh) * 3 + dw], A1[j]);
                A2[j] = fma2(x, w2[dw], A2[j]);
              }
            }
          }
        }
        __syncwarp();
        if (lane0 && !solo) mbar_arrive_relaxed(&ring_empty[r]);
        if (++r == kFrames) { r = 0; rph ^= 1u; }
        if (!first) emit_out(A2);
      };

      int t = 0;
      for (; t + 3 <= p.T; t += 3) {
        step(t == 0, acc[1], acc[0], acc[2]);
        step(false, acc[2], acc[1], acc[0]);
        step(false, acc[0], acc[2], acc[1]);
      }
      // tail (T % 3 frames), then the last output frame, which never sees a dt=2 contribution
      // (temporal zero padding).  The loop above returns the three sets to their original roles, so
      // the accumulators stay in fixed registers (no rotation moves).
      const int rem = p.T - t;
      if (rem >= 1) step(t == 0, acc[1], acc[0], acc[2]);
      if (rem == 2) step(false, acc[2], acc[1], acc[0]);
      if (rem == 0) emit_out(acc[2]);
      else if (rem == 1) emit_out(acc[0]);
      else emit_out(acc[1]);

      if (p.partial != nullptr) {
        float* red = s_red + item_par * (kMaxQ * CH);
        if (in_slot) {
          red[slot * CH + 2 * cp] = ssum.x;
          red[slot * CH + 2 * cp + 1] = ssum.y;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kStencilThreads) : "memory");
        for (int ch = st; ch < CH; ch += kStencilThreads) {
          if (c0 + ch < p.Cs) {
            float a = 0.f;
            for (int k = 0; k < p.Q; ++k) a += red[k * CH + ch];
            p.partial[(static_cast<long>(n) * p.tiles + tile) * p.Cs + c0 + ch] = a;
          }
        }
      }
    }
  }

  tcgen05_before_sync();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------ host
struct Plan {
  int CH, SW, Q, tiles_w, tiles_h, chunks, BH, BW, MB, KC, k16, CHN, XS, NA, acc_cols;
  int x_kc_bytes, x_stage_bytes, x_box_bytes, slot_bytes, stage_bytes;
  int off_wa, off_x, off_ring, off_stage, off_red, off_w2, off_bias;
  size_t smem;
  bool ok;
};

static int ring_pitch(int CH) { return ((CH * 4 / 16) & 1) ? CH * 4 : CH * 4 + 16; }

// The stencil warps are the bottleneck and every work item costs them T steps of SW columns, so the
// plan minimises (items per clip) x (SW + per-step overhead), subject to: 384 stencil threads
// (Q rows x CH/2 channel pairs), at least two TMEM accumulators of MB x CHN columns, and the shared
// memory of one CTA per SM.
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
      int qmax = kStencilThreads / (CH / 2);
      if (qmax > kMaxQ) qmax = kMaxQ;
      if (qmax > Ho) qmax = Ho;
      for (int Q = qmax; Q >= 1; --Q) {
        for (int XS = kMaxXS; XS >= 2; --XS) {
          Plan pl{};
          pl.CH = CH; pl.SW = SW; pl.Q = Q; pl.XS = XS;
          pl.chunks = (Cs + CH - 1) / CH;
          pl.BW = (SW - 1) * stride + 3;
          pl.BH = (Q - 1) * stride + 3;
          pl.MB = (pl.BH * pl.BW + 127) / 128;
          pl.CHN = (CH + 15) / 16 * 16;
          pl.acc_cols = (pl.MB * pl.CHN + 31) / 32 * 32;
          pl.NA = 512 / pl.acc_cols;
          if (pl.NA > kMaxAcc) pl.NA = kMaxAcc;
          if (pl.NA < 2) continue;
          pl.k16 = (Cin + 15) / 16;
          pl.KC = (pl.k16 + 3) / 4;
          // whole 8-row swizzle atoms; the MMA of the last 128-pixel block reads on into whatever
          // follows (rows >= npix only produce accumulator rows nobody drains) -- checked below to
          // stay inside the allocation
          pl.x_kc_bytes = (pl.BH * pl.BW + 7) / 8 * 1024;
          pl.x_stage_bytes = pl.KC * pl.x_kc_bytes;
          pl.x_box_bytes = pl.BH * pl.BW * 128;
          pl.slot_bytes = (pl.BH * pl.BW * ring_pitch(CH) + 127) / 128 * 128;
          pl.stage_bytes = (Q * SW * CH * 2 + 127) / 128 * 128;
          int off = 512;                                   // barriers + tmem slot
          off = (off + 1023) / 1024 * 1024;
          pl.off_wa = off;   off += pl.KC * pl.CHN * 128;
          off = (off + 1023) / 1024 * 1024;
          pl.off_x = off;    off += XS * pl.x_stage_bytes;
          pl.off_ring = off; off += kFrames * pl.slot_bytes;
          pl.off_stage = off; off += kOut * pl.stage_bytes;
          pl.off_red = off;  off += 2 * kMaxQ * CH * 4;
          pl.off_w2 = off;   off += 9 * CH * 4;
          off = (off + 1023) / 1024 * 1024;
          pl.off_bias = off; off += 1024 + pl.CHN * 128;   // ones atom + bias operand tile
          const int mma_end = pl.off_x + XS * pl.x_stage_bytes - pl.x_kc_bytes + pl.MB * 16384;
          if (off < mma_end) off = mma_end;
          pl.smem = (size_t)off + 1024;                    // + alignment slack
          if ((int)pl.smem > max_smem) continue;
          if (pl.smem < 120 * 1024) pl.smem = 120 * 1024;  // one CTA per SM: it owns the whole TMEM
          pl.tiles_w = (Wo + SW - 1) / SW;
          pl.tiles_h = (Ho + Q - 1) / Q;
          const double items = (double)pl.tiles_w * pl.tiles_h * pl.chunks;
          const double drain = 0.004 * pl.MB * 128 * pl.CHN / 64.0;      // relative cost of the drain per step
          // per item: T steps of SW columns + a fixed part per step; an x ring of only two stages
          // leaves the TMA latency exposed (measured ~10 %)
          double cost = items * (SW + 2.0 + drain) * (XS >= 3 ? 1.0 : 1.1);
          pl.ok = true;
          if (cost < best_cost) { best_cost = cost; best = pl; }
          break;                                           // deepest x ring that fits
        }
      }
    }
  }
  return best;
}

template <int S, int SW, int CH>
static int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& to, const Params& p,
                  const Plan& pl, int N, cudaStream_t st) {
  auto kern = ab_persist_kernel<S, SW, CH>;
  static SmemOptIn optin;                              // one per template instance, per device inside
  const cudaError_t e = ensure_dynamic_smem(kern, optin, pl.smem);
  if (e != cudaSuccess) {
    set_error("x3d_expand_dw2_fwd: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e));
    return X3D_ERR_LAUNCH;
  }
  const long total = (long)N * pl.tiles_w * pl.tiles_h;
  long gx = device_sm_count() / pl.chunks;
  if (gx < 1) gx = 1;
  if (gx > total) gx = total;
  dim3 grid((unsigned)gx, pl.chunks);
  kern<<<grid, kThreads, pl.smem, st>>>(tx, tw, to, p);
  return check_launch("x3d_expand_dw2_fwd");
}

template <int S>
static int dispatch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& to, const Params& p,
                    const Plan& pl, int N, cudaStream_t st) {
#define X3D_ABP(SWW, CHH) \
  if (pl.SW == SWW && pl.CH == CHH) return launch<S, SWW, CHH>(tx, tw, to, p, pl, N, st)
  X3D_ABP(8, 56); X3D_ABP(8, 64); X3D_ABP(8, 72);
  X3D_ABP(7, 56); X3D_ABP(7, 64); X3D_ABP(7, 72);
  if constexpr (S == 2) { X3D_ABP(4, 56); X3D_ABP(4, 64); X3D_ABP(4, 72); }
#undef X3D_ABP
  set_error("x3d_expand_dw2_fwd: no kernel for SW=%d CH=%d", pl.SW, pl.CH);
  return X3D_ERR_UNSUPPORTED;
}

}  // namespace abp
}  // namespace x3d

using namespace x3d;

extern "C" int x3d_expand_dw2_partial_blocks(int T, int H, int W, int Cin, int C, int stride) {
  if (T <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || Cin <= 0 || Cin % 8 || (stride != 1 && stride != 2)) return 0;
  const int ms = device_max_smem();
  if (ms <= 0) return 0;
  const abp::Plan pl = abp::make_plan(H, W, Cin, C, stride, ms);
  return pl.ok ? pl.tiles_w * pl.tiles_h : 0;
}

extern "C" int x3d_expand_dw2_fwd(const void* x, const void* wa, const float* bias_a, const float* wb,
                                  const float* bias_b, void* out, float* se_partial, int N, int T, int H,
                                  int W, int Cin, int C, int Kpad, int Npad, int stride, int pad_h,
                                  int pad_w, int act, void* stream) {
  X3D_REQUIRE(x && wa && bias_a && wb && bias_b && out, X3D_ERR_INVALID_ARG, "x3d_expand_dw2_fwd: null pointer");
  X3D_REQUIRE(act == 0 || (act == 1 && se_partial == nullptr), X3D_ERR_INVALID_ARG,
              "x3d_expand_dw2_fwd: act=%d (0, or 1 = swish without SE sums: with SE the scale comes first)", act);
  X3D_REQUIRE(C > 0 && C % 8 == 0 && Cin > 0 && Cin % 8 == 0, X3D_ERR_INVALID_ARG,
              "x3d_expand_dw2_fwd: Cin=%d / C=%d must be multiples of 8", Cin, C);
  X3D_REQUIRE(Kpad % 64 == 0 && Kpad >= Cin && Npad % 16 == 0 && Npad >= C, X3D_ERR_INVALID_ARG,
              "x3d_expand_dw2_fwd: bad packed weight extents Kpad=%d Npad=%d", Kpad, Npad);
  X3D_REQUIRE(stride == 1 || stride == 2, X3D_ERR_UNSUPPORTED, "x3d_expand_dw2_fwd: stride %d", stride);
  X3D_REQUIRE(N > 0 && T > 0 && H > 0 && W > 0, X3D_ERR_INVALID_ARG, "x3d_expand_dw2_fwd: bad extent");
  X3D_REQUIRE(pad_h >= 0 && pad_h <= 1 && pad_w >= 0 && pad_w <= 1, X3D_ERR_INVALID_ARG, "x3d_expand_dw2_fwd: pad_before must be 0 or 1");
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(wa) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_expand_dw2_fwd: pointers must be 16-byte aligned");
  X3D_REQUIRE(device_sm_count() > 0 && device_is_sm100(), X3D_ERR_NO_DEVICE, "x3d_expand_dw2_fwd: needs an sm_100 device");
  EncodeTiledFn enc = tensor_map_encoder();
  X3D_REQUIRE(enc != nullptr, X3D_ERR_NO_DEVICE, "x3d_expand_dw2_fwd: cuTensorMapEncodeTiled unavailable");
  const abp::Plan pl = abp::make_plan(H, W, Cin, C, stride, device_max_smem());
  X3D_REQUIRE(pl.ok, X3D_ERR_UNSUPPORTED, "x3d_expand_dw2_fwd: no tile plan for H=%d W=%d Cin=%d C=%d stride=%d", H, W, Cin, C, stride);
  X3D_REQUIRE(pl.chunks <= 65535, X3D_ERR_UNSUPPORTED, "x3d_expand_dw2_fwd: too many channel chunks");
  X3D_REQUIRE((long)N * pl.tiles_w * pl.tiles_h < (1L << 31), X3D_ERR_UNSUPPORTED, "x3d_expand_dw2_fwd: too many work items");

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
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_expand_dw2_fwd: input tensor map failed (%d) for [%d,%d,%d,%d,%d] box [64,%d,%d]",
                (int)r, N, T, H, W, Cin, pl.BW, pl.BH);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Npad};
    cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)pl.CHN};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wa), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_expand_dw2_fwd: weight tensor map failed (%d)", (int)r);
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)Wo * C * 2, (cuuint64_t)Ho * Wo * C * 2,
                             (cuuint64_t)T * Ho * Wo * C * 2};
    cuuint32_t box[5] = {(cuuint32_t)pl.CH, (cuuint32_t)pl.SW, (cuuint32_t)pl.Q, 1, 1};
    CUresult r = enc(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, out, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_expand_dw2_fwd: output tensor map failed (%d)", (int)r);
  }

  abp::Params p;
  p.bias_a = bias_a; p.w = wb; p.bias = bias_b; p.partial = se_partial;
  p.N = N; p.T = T; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.Cs = C;
  p.Q = pl.Q; p.tiles_w = pl.tiles_w; p.tiles = pl.tiles_w * pl.tiles_h;
  p.pad_h = pad_h; p.pad_w = pad_w;
  p.KC = pl.KC; p.k16 = pl.k16; p.MB = pl.MB; p.CHN = pl.CHN; p.XS = pl.XS; p.NA = pl.NA; p.acc_cols = pl.acc_cols;
  p.x_kc_bytes = pl.x_kc_bytes; p.x_stage_bytes = pl.x_stage_bytes; p.x_box_bytes = pl.x_box_bytes;
  p.slot_bytes = pl.slot_bytes; p.stage_bytes = pl.stage_bytes;
  p.off_wa = pl.off_wa; p.off_x = pl.off_x; p.off_ring = pl.off_ring; p.off_stage = pl.off_stage;
  p.off_red = pl.off_red; p.off_w2 = pl.off_w2; p.off_bias = pl.off_bias;
  p.act = act;
  { const char* e = getenv("X3D_ABP_DEBUG"); p.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("X3D_ABP_SLEEP"); p.sleep_ns = e ? (uint32_t)atoi(e) : 4000u; }
  { const char* e = getenv("X3D_ABP_NA"); if (e && atoi(e) >= 2 && atoi(e) < p.NA) p.NA = atoi(e); }
  if (p.dbg & 16)
    fprintf(stderr, "abp plan H=%d W=%d Cin=%d C=%d s=%d: CH=%d SW=%d Q=%d XS=%d NA=%d MB=%d KC=%d chunks=%d tiles=%dx%d smem=%zu\n",
            H, W, Cin, C, stride, pl.CH, pl.SW, pl.Q, pl.XS, p.NA, pl.MB, pl.KC, pl.chunks, pl.tiles_h, pl.tiles_w, pl.smem);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return stride == 1 ? abp::dispatch<1>(tx, tw, to, p, pl, N, st) : abp::dispatch<2>(tx, tw, to, p, pl, N, st);
}
