// Pointwise (1x1x1) convolution on the Blackwell tensor cores: bf16 operands, fp32 accumulate in
// TMEM.  D[M, Nc] = act(bias + pro(A[M, K]) * W[K, Nc] (+ R)).
//
//   * persistent CTAs (grid.x <= #SMs x n-tiles), one CTA per SM; the CTA's weight slice
//     W[n0:n0+NT, 0:Kpad] is fetched ONCE by TMA and stays in shared memory (128B-swizzled,
//     K-major) for the kernel's life;
//   * A streams through a ring of 128x64 bf16 stages filled by TMA (cp.async.bulk.tensor, 128B
//     swizzle, out-of-bounds rows/channels zero-filled, so K and M need no padding in HBM);
//   * one elected thread issues tcgen05.mma (M=128, N=NT, K=16) into a double-buffered TMEM
//     accumulator, tcgen05.commit releases ring stages / publishes the accumulator via mbarriers;
//   * 2 groups of 4 epilogue warps (one per TMEM accumulator) read TMEM with tcgen05.ld, add bias
//     (+ residual, prefetched), ReLU, pack bf16 into swizzled staging and TMA-store full lines;
//   * optional prologue (projection conv, model.py:311-317): 8 transform warps apply
//     swish(se[clip,k] * a) in place on each landed stage before the MMA consumes it;
//   * optional column means (conv_5 + pool_5, model.py:117-118): the epilogue also reduces its staged
//     bf16 tile over rows, 64 at a time, so the global average pool never reads the conv output back
//     (and with store_d = 0 the output is not even written);
//   * optional second A source (ResBlock's strided shortcut conv + bn_r, model.py:360-367,386-389):
//     the K loop continues over the block INPUT, read through a 4-D tensor map whose strides pick
//     the pixels (t, s*ho, s*wo), against the shortcut's weights stacked under the projection's:
//     D = bn_c(c(h)) + bn_r(r(x)) in one fp32 accumulator, no gather kernel, no residual tensor.
//
// Reference call sites replaced: Bottleneck.a/bn_a/relu (model.py:306-308), Bottleneck.c/bn_c +
// ResBlock add/relu (model.py:317-318,389-392), conv5 (model.py:117).
#include <stdlib.h>

#include "tma_common.cuh"

namespace x3d {
namespace tc {

constexpr int kBlockM = 128, kBlockK = 64, kStageBytes = kBlockM * kBlockK * 2;
using namespace ptx;

// Timeline trace of one epilogue group (diagnostics only: nvcc -DX3D_PW_TRACE; tools/trace_pw.py).
#ifdef X3D_PW_TRACE
__device__ long long g_pw_trace[8 * 256];
#define X3D_TRACE(slot)                                                                  \
  do {                                                                                   \
    if (trace_on && trace_n < 256) g_pw_trace[trace_n * 8 + (slot)] = clock64();         \
  } while (0)
#else
#define X3D_TRACE(slot) do {} while (0)
#endif

__device__ __forceinline__ void tcgen05_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows are 128 B apart inside
// an 8-row (1024 B) swizzle atom, atoms stacked every 1024 B (SBO); LBO unused; version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= static_cast<uint64_t>(0) << 16;                           // LBO            [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // SBO            [32,46)
  d |= static_cast<uint64_t>(1) << 46;                           // version = 1    [46,48)
  d |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B   [61,64)
  return d;
}
// Instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n.
__device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(kBlockM >> 4) << 24);
}

struct Params {
  const float* bias; const bf16* R; const float* se; bf16* D;
  long M; long rows_per_clip;
  int Kc;          // stored input channels (multiple of 8)
  int Nc;          // stored output channels (multiple of 8)
  int ldr, ldd;
  int NT;          // N tile of this launch (multiple of 16, <= 256)
  int KC;          // number of 64-wide K chunks (both sources)
  int k16_last;    // number of K=16 MMAs in the last chunk of the LAST source (1..4)
  int KC1;         // chunks of the first source (== KC without a second source)
  int k16_last1;   // K=16 MMAs in the last chunk of the first source
  int a2_ppf;      // second source: output pixels per frame (Ho*Wo), 0 without one
  int a2_wo;       // second source: output row length
  float* colmean;  // [2*tiles, ldc] means over the 64-row halves of every tile (conv5 + pool), or null
  int ldc;
  int store_d;     // 0: D is not written (only the column means are wanted)
  int stages;      // A ring depth
  int tmem_cols;   // power of two >= 2*NT
  int relu, swish;
  int contiguous;  // tile assignment (see the kernel)
};

constexpr int kEpiWarps = 8, kProWarps = 8, kOutBufs = 4;
// transform groups (each takes every kProGroups-th stage).  The A ring must be at least this deep: a group's
// FIRST wait has to be on the ring's first pass (a parity wait cannot tell pass 1 from 'never filled').
// Measured (stage 2 / 3 / 4 projection with SE, 80 clips): 1 group 0.267 / 0.130 / 0.078 ms, 2 groups
// 0.234 / 0.122 / 0.076, 4 groups 0.229 / 0.121 / 0.073 (and needs 4 stages, which K = 432 does not leave).
constexpr int kProGroups = 2;
constexpr int kProRows = 8 * kProWarps / kProGroups / 2;   // row stride of a transform thread: (threads per group) / 8   // 2 epilogue groups x 2 staging buffers
constexpr int kThreadsPlain = 64 + 32 * kEpiWarps, kThreadsPro = kThreadsPlain + 32 * kProWarps;

template <bool kPro>
__global__ void __launch_bounds__(kPro ? kThreadsPro : kThreadsPlain, 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmA2,
             const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.NT;
  const int w_chunk_bytes = p.NT * 128;

  uint8_t* sW = smem;
  uint8_t* sA = sW + p.KC * w_chunk_bytes;
  uint8_t* sD = sA + p.stages * kStageBytes;          // [kOutBufs] 128x64 bf16 output staging, 128B swizzle
  float* sBias = reinterpret_cast<float*>(sD + kOutBufs * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 256);
  uint64_t* full = bars;                      // [stages]  TMA landed
  uint64_t* empty = full + p.stages;          // [stages]  MMA done reading
  uint64_t* xform = empty + p.stages;         // [stages]  prologue applied
  uint64_t* w_full = xform + p.stages;        // weights landed
  uint64_t* t_full = w_full + 1;              // [2] accumulator ready
  uint64_t* t_empty = t_full + 2;             // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  for (int i = threadIdx.x; i < p.NT; i += blockDim.x) {
    const int col = n0 + i;
    sBias[i] = (p.bias != nullptr && col < p.Nc) ? p.bias[col] : 0.f;
  }
  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmD);
    if (p.KC > p.KC1) prefetch_tmap(&tmA2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&xform[s], 32 * kProWarps / kProGroups);   // one transform group per stage
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 16 * kEpiWarps);          // one epilogue group (4 warps) per buffer
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // Tiles of this CTA: tile_lo, tile_lo + t_stride, ... < tile_hi.  Plain GEMMs interleave the CTAs
  // (t_stride = grid: the 148 CTAs sweep one window of memory together); with the SE prologue a CTA takes
  // one contiguous range, so it stays inside one or two clips and their scale vectors stay cached.
  const long num_tiles = (p.M + kBlockM - 1) / kBlockM;
  long tile_lo = blockIdx.x, tile_hi = num_tiles, t_stride = gridDim.x;
  if (p.contiguous) {
    const long t_per = num_tiles / gridDim.x, t_rem = num_tiles % gridDim.x;
    tile_lo = blockIdx.x * t_per + (blockIdx.x < t_rem ? blockIdx.x : t_rem);
    tile_hi = tile_lo + t_per + (blockIdx.x < t_rem ? 1 : 0);
    t_stride = 1;
  }

  // Programmatic dependent launch: everything above (and the weight fetch below) touches nothing the
  // previous kernel of the stream writes; activations are only accessed after pdl_wait().
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(w_full, static_cast<uint32_t>(p.KC * w_chunk_bytes));
      for (int kc = 0; kc < p.KC; ++kc)
        tma_load_2d(sW + kc * w_chunk_bytes, &tmW, kc * kBlockK, n0, w_full);
      pdl_wait();
      pdl_trigger();
      int s = 0;
      uint32_t ph = 0;
      for (long tile = tile_lo; tile < tile_hi; tile += t_stride) {
        const int row0 = static_cast<int>(tile * kBlockM);
        // second source: the tile's 128 output pixels are whole rows of one frame (or whole frames)
        int nt0 = 0, ho0 = 0;
        if (p.a2_ppf) {
          nt0 = row0 / p.a2_ppf;
          ho0 = (row0 - nt0 * p.a2_ppf) / p.a2_wo;
        }
        for (int kc = 0; kc < p.KC; ++kc) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kStageBytes);
          if (kc < p.KC1) tma_load_2d(sA + s * kStageBytes, &tmA, kc * kBlockK, row0, &full[s]);
          else if (p.a2_ppf) tma_load_4d(smem_u32(sA + s * kStageBytes), &tmA2, (kc - p.KC1) * kBlockK, 0, ho0, nt0, &full[s]);
          else tma_load_2d(sA + s * kStageBytes, &tmA2, (kc - p.KC1) * kBlockK, row0, &full[s]);   // dense [M, K2] rows
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    pdl_trigger();                                // (reads only what the producer loads after its own wait)
    const uint32_t idesc = make_idesc_bf16(p.NT);
    mbar_wait(w_full, 0);
    int s = 0;
    uint32_t ph = 0;
    long it = 0;
    for (long tile = tile_lo; tile < tile_hi; tile += t_stride, ++it) {
      const int as = static_cast<int>(it & 1);
      const uint32_t aph = static_cast<uint32_t>((it >> 1) & 1);
      mbar_wait(&t_empty[as], aph ^ 1);
      tcgen05_after_sync();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.NT);
      for (int kc = 0; kc < p.KC; ++kc) {
        mbar_wait(kPro ? &xform[s] : &full[s], ph);
        tcgen05_after_sync();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(sA + s * kStageBytes);
          const uint32_t b_addr = smem_u32(sW + kc * w_chunk_bytes);
          const int nk = (kc == p.KC - 1) ? p.k16_last : (kc == p.KC1 - 1 ? p.k16_last1 : 4);
          for (int k = 0; k < nk; ++k)
            umma_bf16(d_tmem, make_desc_sw128(a_addr + k * 32), make_desc_sw128(b_addr + k * 32),
                      idesc, (kc | k) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);                       // frees the stage when these MMAs retire
          if (kc == p.KC - 1) umma_commit(&t_full[as]);  // accumulator complete
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ------------------------------------------------------------------ epilogue (2 groups x 4 warps)
    // Group g owns TMEM accumulator g and every second tile, so two epilogues are in flight; inside
    // a group warp w reads TMEM lane quarter w%4.  Results are staged in 128B-swizzled shared
    // memory and leave as full lines through TMA box stores (clipped at M / Nc).
    pdl_wait();                                   // residual rows / output stores
    pdl_trigger();
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;
    const bool leader = q == 2 && lane == 0;      // first warp of each group is warp 2 / warp 6
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    const int r_loc = q * 32 + lane;              // row inside the tile
    const uint32_t sw_row = static_cast<uint32_t>(r_loc) * 128, sw_x = static_cast<uint32_t>(r_loc & 7);
    const int n_sub = (p.NT + 63) >> 6;
    const uint32_t stage0 = smem_u32(sD + grp * 2 * kStageBytes);
    const uint32_t bias_s = smem_u32(sBias);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                           static_cast<uint32_t>(grp * p.NT);
    uint32_t sub_ctr = 0, aph = 0;
#ifdef X3D_PW_TRACE
    const bool trace_on = blockIdx.x == 7 && blockIdx.y == 0 && grp == 0 && leader;
    int trace_n = 0;
#endif
    for (long tile = tile_lo + grp * t_stride; tile < tile_hi; tile += 2 * t_stride, aph ^= 1) {
      X3D_TRACE(0);
      const long row = tile * kBlockM + r_loc;
      const bool row_ok = row < p.M;
      const bf16* rrow = (p.R && row_ok) ? p.R + row * p.ldr + n0 : nullptr;
      // residual of the first sub-tile is requested before the accumulator is waited for
      uint4 rr[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        rr[c] = (rrow && c * 8 < p.NT && n0 + c * 8 < p.Nc) ? __ldg(reinterpret_cast<const uint4*>(rrow + c * 8))
                                                           : make_uint4(0, 0, 0, 0);
      // The residual rows come from DRAM (several kernels ago) and their latency is exposed in the
      // drain below (timeline: tools/trace_pw.py): pull the NEXT tile's rows of this group into L2 now
      if (p.R) {
        const long nrow = row + 2L * t_stride * kBlockM;
        if (nrow < p.M && tile + 2 * t_stride < tile_hi) {
          const char* np_ = reinterpret_cast<const char*>(p.R + nrow * p.ldr + n0);
          for (int off = 0; off < p.NT * 2; off += 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + off));
        }
      }
      mbar_wait(&t_full[grp], aph);
      X3D_TRACE(1);
      tcgen05_after_sync();
      for (int sb = 0; sb < n_sub; ++sb, ++sub_ctr) {
        if (n0 + sb * 64 >= p.Nc) break;          // uniform across the CTA
        const uint32_t buf = stage0 + (sub_ctr & 1) * kStageBytes + sw_row;
        if (sb > 0) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int cc = sb * 64 + c * 8;
            rr[c] = (rrow && cc < p.NT && n0 + cc < p.Nc) ? __ldg(reinterpret_cast<const uint4*>(rrow + cc))
                                                         : make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c0 = sb * 64 + g * 16;        // column inside the N tile
          if (c0 >= p.NT || n0 + c0 >= p.Nc) continue;
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
#ifdef X3D_PW_TRACE
          if (g == 0 && sb == 0 && trace_on && trace_n < 256) g_pw_trace[trace_n * 8 + 7] = clock64();
#endif
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cc = c0 + h * 8;
            float4 b0, b1;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0.x), "=f"(b0.y), "=f"(b0.z), "=f"(b0.w) : "r"(bias_s + cc * 4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b1.x), "=f"(b1.y), "=f"(b1.z), "=f"(b1.w) : "r"(bias_s + cc * 4 + 16));
            float2 y[4];
            y[0] = __fadd2_rn(make_float2(__uint_as_float(v[h * 8 + 0]), __uint_as_float(v[h * 8 + 1])), make_float2(b0.x, b0.y));
            y[1] = __fadd2_rn(make_float2(__uint_as_float(v[h * 8 + 2]), __uint_as_float(v[h * 8 + 3])), make_float2(b0.z, b0.w));
            y[2] = __fadd2_rn(make_float2(__uint_as_float(v[h * 8 + 4]), __uint_as_float(v[h * 8 + 5])), make_float2(b1.x, b1.y));
            y[3] = __fadd2_rn(make_float2(__uint_as_float(v[h * 8 + 6]), __uint_as_float(v[h * 8 + 7])), make_float2(b1.z, b1.w));
            if (p.R) {
              const uint4 ru = rr[g * 2 + h];
              const uint32_t rw[4] = {ru.x, ru.y, ru.z, ru.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
                y[j] = __fadd2_rn(y[j], make_float2(__uint_as_float(rw[j] << 16),
                                                    __uint_as_float(rw[j] & 0xffff0000u)));
            }
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 t = __float22bfloat162_rn(y[j]);
              if (p.relu) t = __hmax2(t, zero2);
              o[j] = *reinterpret_cast<uint32_t*>(&t);
            }
            const uint32_t chunk = static_cast<uint32_t>(g * 2 + h);          // 16-byte chunk 0..7
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf + ((chunk ^ sw_x) << 4)),
                         "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
        }
        X3D_TRACE(2);
        if (sb == n_sub - 1 || n0 + (sb + 1) * 64 >= p.Nc) {
          tcgen05_before_sync();                  // last TMEM read of this tile is done
          mbar_arrive(&t_empty[grp]);
        }
        fence_proxy_async();
        X3D_TRACE(3);
        // the store issued after the previous barrier has (long since) finished reading its
        // staging buffer: checked here, off the critical path, so that buffer is free again for
        // everybody once this barrier is passed
        if (leader) tma_store_wait_read<0>();
        X3D_TRACE(4);
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        X3D_TRACE(5);
        if (leader && p.store_d) {
          tma_store_2d(&tmD, stage0 + (sub_ctr & 1) * kStageBytes, n0 + sb * 64,
                       static_cast<int>(tile * kBlockM));
          tma_store_commit();
        }
        if (p.colmean != nullptr) {
          // column means of the staged (bf16-rounded) tile: thread = (column, 64-row half); a warp
          // reads 32 neighbouring columns of one row per load (one 64-byte piece of a swizzled line)
          const int ccol = r_loc & 63, half = r_loc >> 6;
          const int col_g = n0 + sb * 64 + ccol;
          const long row_base = tile * kBlockM + half * 64;
          int nrows = p.M - row_base < 64 ? static_cast<int>(p.M - row_base) : 64;
          if (ccol < p.NT - sb * 64 && col_g < p.Nc && nrows > 0) {
            const uint32_t sbuf = stage0 + (sub_ctr & 1) * kStageBytes;
            float sum = 0.f;
            for (int r = 0; r < nrows; ++r) {
              const int row = half * 64 + r;
              unsigned short h;
              asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h)
                           : "r"(sbuf + row * 128 + ((((ccol >> 3) ^ (row & 7)) << 4) | ((ccol & 7) << 1))));
              sum += __uint_as_float(static_cast<uint32_t>(h) << 16);
            }
            p.colmean[(tile * 2 + half) * p.ldc + col_g] = sum * (1.f / 64.f);
          }
        }
        X3D_TRACE(6);
#ifdef X3D_PW_TRACE
        if (trace_on) ++trace_n;
#endif
      }
    }
    if (leader) tma_store_wait_read<0>();
  } else if (kPro) {
    // ------------------------------------------------------------------ prologue transform (2 x 4 warps)
    // In place on the landed stage: a <- swish(se[clip, k] * a).  The two groups of four warps take
    // alternate stages, so one group's load -> MUFU -> store -> proxy fence -> arrive chain overlaps
    // the other's (with all eight warps on one stage that chain was the kernel's critical path).
    // A thread owns one physical 16-byte chunk column (8 channels) of rows rbase + 16*i; 16 is a
    // multiple of the swizzle period, so its channel offset k is the same for all 8 rows and the 8 SE
    // factors are fetched once per stage -- requested one stage of the group ahead (the L2 round trip
    // was 17 % of these warps' time); tiles that straddle two clips take the per-row path.
    pdl_wait();                                   // the SE factors come from the previous kernel
    pdl_trigger();
    const int tt = threadIdx.x - kThreadsPlain;   // 0..255
    constexpr int kGT = 32 * kProWarps / kProGroups;          // threads per group
    const int grp = tt / kGT;                     // the stages j with j % kProGroups == grp are this thread's
    const int pchunk = tt & 7;                    // physical 16-byte chunk inside the 128-byte row
    const int rbase = (tt % kGT) >> 3;            // 0 .. kGT/8-1 (a multiple of 8 rows apart: same swizzle phase)
    const int kin = (pchunk ^ (rbase & 7)) << 3;  // channel offset inside the 64-wide chunk
    const uint32_t rpc = static_cast<uint32_t>(p.rows_per_clip > 0 ? p.rows_per_clip : 0x7fffffff);
    const uint32_t step = static_cast<uint32_t>(t_stride) * kBlockM;
    // position of a stage: (tile, K chunk) and (clip, row inside the clip) of the tile's first row,
    // advanced without divisions
    struct Pos { long tile; int kc; uint32_t clip, rem; };
    auto advance = [&](Pos& q) {
      if (++q.kc == p.KC) {
        q.kc = 0;
        q.tile += t_stride;
        q.rem += step;
        while (q.rem >= rpc) { q.rem -= rpc; ++q.clip; }
      }
    };
    auto se_request = [&](const Pos& q, float4& s0, float4& s1) {
      s0 = make_float4(1.f, 1.f, 1.f, 1.f);
      s1 = s0;
      const int k = q.kc * kBlockK + kin;
      if (p.se != nullptr && q.tile < tile_hi && k < p.Kc && q.kc < p.KC1 && q.rem + kBlockM <= rpc) {
        const float4* sp = reinterpret_cast<const float4*>(p.se + static_cast<long>(q.clip) * p.Kc + k);
        s0 = __ldg(sp);
        s1 = __ldg(sp + 1);
      }
    };
    Pos cur;
    cur.tile = tile_lo; cur.kc = 0;
    cur.clip = static_cast<uint32_t>(tile_lo * static_cast<long>(kBlockM) / rpc);
    cur.rem = static_cast<uint32_t>(tile_lo * static_cast<long>(kBlockM) - static_cast<long>(cur.clip) * rpc);
    int s = 0;
    uint32_t ph = 0;
    for (int g = 0; g < grp; ++g) {
      advance(cur);
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    float4 nx0, nx1;
    se_request(cur, nx0, nx1);
    while (cur.tile < tile_hi) {
      const uint32_t clip0 = cur.clip, rem0 = cur.rem;
      const bool one_clip = rem0 + kBlockM <= rpc;            // whole tile inside one clip
      const int k = cur.kc * kBlockK + kin;
      // beyond Kc the stage holds TMA zeros; chunks of the second source are not transformed
      const bool k_ok = k < p.Kc && cur.kc < p.KC1;
      float2 sc[4];
      sc[0] = make_float2(nx0.x, nx0.y); sc[1] = make_float2(nx0.z, nx0.w);
      sc[2] = make_float2(nx1.x, nx1.y); sc[3] = make_float2(nx1.z, nx1.w);
#pragma unroll
      for (int g = 0; g < kProGroups; ++g) advance(cur);
      se_request(cur, nx0, nx1);
      mbar_wait(&full[s], ph);
      if (k_ok) {
        const uint32_t addr0 = smem_u32(sA + s * kStageBytes) + rbase * 128 + pchunk * 16;
        float2 hs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) hs[j] = __fmul2_rn(sc[j], make_float2(0.5f, 0.5f));
#pragma unroll
        for (int half = 0; half < kBlockM / (kGT / 8) / 4; ++half) {
          // four rows at a time: their shared loads are issued back to back before any math
          uint32_t w[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w[i][0]), "=r"(w[i][1]), "=r"(w[i][2]), "=r"(w[i][3])
                         : "r"(addr0 + (half * 4 + i) * (kGT / 8 * 128)));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (p.se && !one_clip) {
              uint32_t x = rem0 + rbase + (kGT / 8) * (half * 4 + i), clip = clip0;
              while (x >= rpc) { x -= rpc; ++clip; }
              if (static_cast<long>(clip) * rpc + x < p.M) {
                const float4* sp = reinterpret_cast<const float4*>(p.se + static_cast<long>(clip) * p.Kc + k);
                const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
                sc[0] = make_float2(s0.x, s0.y); sc[1] = make_float2(s0.z, s0.w);
                sc[2] = make_float2(s1.x, s1.y); sc[3] = make_float2(s1.z, s1.w);
#pragma unroll
                for (int j = 0; j < 4; ++j) hs[j] = __fmul2_rn(sc[j], make_float2(0.5f, 0.5f));
              }
            }
            if (p.swish) {
              // swish(s x) = h + h tanh(h) with h = (s/2) x: one multiply, two MUFU, one FMA per pair
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 h = __fmul2_rn(make_float2(__uint_as_float(w[i][j] << 16), __uint_as_float(w[i][j] & 0xffff0000u)), hs[j]);
                float2 t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
                __nv_bfloat162 t2 = __float22bfloat162_rn(__ffma2_rn(h, t, h));
                w[i][j] = *reinterpret_cast<uint32_t*>(&t2);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 scj = make_float2(2.f * hs[j].x, 2.f * hs[j].y);
                __nv_bfloat162 t2 = __float22bfloat162_rn(__fmul2_rn(
                    make_float2(__uint_as_float(w[i][j] << 16), __uint_as_float(w[i][j] & 0xffff0000u)), scj));
                w[i][j] = *reinterpret_cast<uint32_t*>(&t2);
              }
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr0 + (half * 4 + i) * (kGT / 8 * 128)), "r"(w[i][0]), "r"(w[i][1]), "r"(w[i][2]), "r"(w[i][3]) : "memory");
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&xform[s]);
      s += kProGroups;
      while (s >= p.stages) { s -= p.stages; ph ^= 1; }
    }
  }

  tcgen05_before_sync();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

// ------------------------------------------------------------------------------ host
static bool make_map_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Strided pixel sampler of the shortcut conv: [C, Wo, Ho, NT] over an NDHWC tensor [NT, Hi, Wi, C]
// with element (c, wo, ho, nt) = in[nt, s*ho, s*wo, c]; box = 64 channels x 128 output pixels.
static bool make_map_sampler(CUtensorMap* m, const void* base, int C, int Hi, int Wi, long NT, int s,
                             int bw, int bh, int bf) {
  EncodeTiledFn enc = tensor_map_encoder();
  if (!enc) return false;
  const int Ho = (Hi - 1) / s + 1, Wo = (Wi - 1) / s + 1;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)NT};
  cuuint64_t strides[3] = {(cuuint64_t)s * C * 2, (cuuint64_t)s * Wi * C * 2, (cuuint64_t)Hi * Wi * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bf};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 128 consecutive output pixels must be whole rows of one frame, or whole frames
static bool sampler_box(int Ho, int Wo, int* bh, int* bf) {
  if (Wo <= 0 || Wo > kBlockM || kBlockM % Wo) return false;
  const int rh = kBlockM / Wo;
  if (rh <= Ho) { if (Ho % rh) return false; *bh = rh; *bf = 1; return true; }
  if (rh % Ho) return false;
  *bh = Ho; *bf = rh / Ho;
  return true;
}

}  // namespace tc
}  // namespace x3d

using namespace x3d;

extern "C" int x3d_pw_tc_sampler_supported(int Hi, int Wi, int stride) {
  if (Hi <= 0 || Wi <= 0 || stride < 1) return 0;
  int bh, bf;
  return tc::sampler_box((Hi - 1) / stride + 1, (Wi - 1) / stride + 1, &bh, &bf) ? 1 : 0;
}

extern "C" int x3d_pw_tc_fwd(const x3d_pw_tc_args* a, void* stream) {
  X3D_REQUIRE(a && a->A && a->Wp && a->D, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: null pointer");
  X3D_REQUIRE(a->M > 0 && a->M < (1L << 31), X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: M out of range");
  X3D_REQUIRE(a->K > 0 && a->K % 8 == 0 && a->lda % 8 == 0 && a->lda >= a->K, X3D_ERR_INVALID_ARG,
              "x3d_pw_tc_fwd: K=%d / lda=%d must be multiples of 8", a->K, a->lda);
  X3D_REQUIRE(a->Nc > 0 && a->Nc % 8 == 0 && a->ldd % 8 == 0 && (!a->R || a->ldr % 8 == 0),
              X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: Nc/ldd/ldr must be multiples of 8");
  X3D_REQUIRE(a->Kpad % 64 == 0 && a->Kpad >= a->K && a->Npad % 16 == 0 && a->Npad >= a->Nc,
              X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: bad packed weight extents Kpad=%d Npad=%d", a->Kpad, a->Npad);
  X3D_REQUIRE(!a->se || a->rows_per_clip > 0, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: rows_per_clip missing");
  X3D_REQUIRE(!a->colmean || (reinterpret_cast<uintptr_t>(a->colmean) & 3) == 0, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: colmean alignment");
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->Wp) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(a->D) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: pointers must be 16-byte aligned");
  X3D_REQUIRE(device_sm_count() > 0 && device_is_sm100(), X3D_ERR_NO_DEVICE, "x3d_pw_tc_fwd: needs an sm_100 device");
  // N tiling: fewest tiles with NT <= 256 whose resident weight slice leaves >= 3 A stages.
  const int k16_total = (a->K + 15) / 16;
  const int KC1 = (k16_total + 3) / 4;                // chunks that actually hold data
  const int k16_last1 = k16_total - (KC1 - 1) * 4;
  int KC = KC1, k16_last = k16_last1, a2_bh = 0, a2_bf = 0, a2_ho = 0, a2_wo = 0;
  if (a->A2) {
    // second source: its K2 channels continue the K loop at column KC1*64 of the packed weight
    X3D_REQUIRE(!a->R, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: A2 (shortcut as extra K) and R (residual) exclude each other");
    X3D_REQUIRE(a->K2 > 0 && a->K2 % 8 == 0 && a->a2_stride >= 0, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: bad second source K2=%d stride=%d", a->K2, a->a2_stride);
    X3D_REQUIRE((reinterpret_cast<uintptr_t>(a->A2) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: A2 must be 16-byte aligned");
    if (a->a2_stride > 0) {
      X3D_REQUIRE(a->a2_hi > 0 && a->a2_wi > 0 && a->a2_nt > 0, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: bad second source %dx%d", a->a2_hi, a->a2_wi);
      a2_ho = (a->a2_hi - 1) / a->a2_stride + 1; a2_wo = (a->a2_wi - 1) / a->a2_stride + 1;
      X3D_REQUIRE(a->M == a->a2_nt * a2_ho * a2_wo, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: M=%ld is not nt*Ho*Wo = %ld*%d*%d",
                  (long)a->M, (long)a->a2_nt, a2_ho, a2_wo);
      X3D_REQUIRE(tc::sampler_box(a2_ho, a2_wo, &a2_bh, &a2_bf), X3D_ERR_UNSUPPORTED,
                  "x3d_pw_tc_fwd: 128-pixel tiles do not align with %dx%d frames (x3d_pw_tc_sampler_supported)", a2_ho, a2_wo);
    }
    const int k16_2 = (a->K2 + 15) / 16, KC2 = (k16_2 + 3) / 4;
    KC = KC1 + KC2; k16_last = k16_2 - (KC2 - 1) * 4;
    X3D_REQUIRE(a->Kpad >= KC * 64, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: Kpad=%d < %d (both sources, 64-wide chunks)", a->Kpad, KC * 64);
  }
  const int budget = device_max_smem() - 1024 /*align*/ - 1024 /*bias*/ - 512 /*barriers*/ -
                     tc::kOutBufs * tc::kStageBytes /*output staging*/;
  // The epilogue stores 64-column boxes, so with more than one N tile the tile width must be a
  // multiple of 64 (a box of one tile must never reach into its neighbour's columns).
  int n_tiles = (a->Npad + 255) / 256;
  int NT = 0, stages = 0;
  for (;; ++n_tiles) {
    const int per = (a->Npad + n_tiles - 1) / n_tiles;
    NT = n_tiles == 1 ? (per + 15) / 16 * 16 : (per + 63) / 64 * 64;
    const int wbytes = KC * NT * 128;
    stages = (budget - wbytes) / tc::kStageBytes;
    if (stages >= 3 || NT <= 64) break;
  }
  X3D_REQUIRE(stages >= 2 && stages >= tc::kProGroups, X3D_ERR_UNSUPPORTED, "x3d_pw_tc_fwd: K=%d too large for shared memory", a->K);
  if (stages > 8) stages = 8;
  n_tiles = (a->Nc + NT - 1) / NT;                     // tiles that hold real columns
  int tmem_cols = 32;
  while (tmem_cols < 2 * NT) tmem_cols *= 2;
  X3D_REQUIRE(tmem_cols <= 512, X3D_ERR_UNSUPPORTED, "x3d_pw_tc_fwd: NT=%d needs too much TMEM", NT);

  CUtensorMap tmA, tmW;
  X3D_REQUIRE(tensor_map_encoder() != nullptr, X3D_ERR_NO_DEVICE, "x3d_pw_tc_fwd: cuTensorMapEncodeTiled unavailable");
  X3D_REQUIRE(tc::make_map_2d(&tmA, a->A, (uint64_t)a->K, (uint64_t)a->M, (uint64_t)a->lda * 2, 64, 128),
              X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: tensor map for A failed (K=%d M=%ld lda=%d)", a->K, (long)a->M, a->lda);
  // The weight box may reach past Npad rows on the last tile: TMA zero-fills.
  X3D_REQUIRE(tc::make_map_2d(&tmW, a->Wp, (uint64_t)a->Kpad, (uint64_t)a->Npad, (uint64_t)a->Kpad * 2, 64, (uint32_t)NT),
              X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: tensor map for W failed (Kpad=%d Npad=%d NT=%d)", a->Kpad, a->Npad, NT);

  CUtensorMap tmD;
  X3D_REQUIRE(tc::make_map_2d(&tmD, a->D, (uint64_t)a->Nc, (uint64_t)a->M, (uint64_t)a->ldd * 2, 64, 128),
              X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: tensor map for D failed (Nc=%d M=%ld ldd=%d)", a->Nc, (long)a->M, a->ldd);
  CUtensorMap tmA2 = tmA;
  if (a->A2 && a->a2_stride > 0)
    X3D_REQUIRE(tc::make_map_sampler(&tmA2, a->A2, a->K2, a->a2_hi, a->a2_wi, (long)a->a2_nt, a->a2_stride, a2_wo, a2_bh, a2_bf),
                X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: tensor map for A2 failed (K2=%d %dx%d stride %d)", a->K2, a->a2_hi, a->a2_wi, a->a2_stride);
  else if (a->A2)
    X3D_REQUIRE(tc::make_map_2d(&tmA2, a->A2, (uint64_t)a->K2, (uint64_t)a->M, (uint64_t)a->K2 * 2, 64, 128),
                X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: tensor map for dense A2 failed (K2=%d M=%ld)", a->K2, (long)a->M);
  tc::Params p;
  p.bias = a->bias; p.R = static_cast<const bf16*>(a->R); p.se = a->se; p.D = static_cast<bf16*>(a->D);
  p.M = a->M; p.rows_per_clip = a->rows_per_clip; p.Kc = a->K; p.Nc = a->Nc; p.ldr = a->ldr; p.ldd = a->ldd;
  p.colmean = a->colmean; p.ldc = a->Nc; p.store_d = a->colmean == nullptr || a->store_d;
  p.KC1 = KC1; p.k16_last1 = k16_last1; p.a2_ppf = (a->A2 && a->a2_stride > 0) ? a2_ho * a2_wo : 0; p.a2_wo = a2_wo;
  p.NT = NT; p.KC = KC; p.k16_last = k16_last; p.stages = stages; p.tmem_cols = tmem_cols;
  p.relu = a->relu; p.swish = a->swish;
  p.contiguous = a->se != nullptr ? 1 : 0;

  const size_t smem = 1024 + (size_t)KC * NT * 128 + (size_t)(stages + tc::kOutBufs) * tc::kStageBytes + 1024 + 512;
  const long num_tiles = (a->M + tc::kBlockM - 1) / tc::kBlockM;
  int gx = device_sm_count() / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > num_tiles) gx = (int)num_tiles;
  dim3 grid(gx, n_tiles);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool pro = (a->se != nullptr) || a->swish;
  cudaError_t e;
  static SmemOptIn optin_pro, optin_plain;             // per device inside (the attribute is per device)
  if (pro) {
    e = ensure_dynamic_smem(tc::pw_tc_kernel<true>, optin_pro, smem, false);
    X3D_REQUIRE(e == cudaSuccess, X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: smem attribute: %s", cudaGetErrorString(e));
    e = launch_pdl(tc::pw_tc_kernel<true>, grid, dim3(tc::kThreadsPro), smem, st, tmA, tmW, tmD, tmA2, p);
    X3D_REQUIRE(e == cudaSuccess, X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: launch: %s", cudaGetErrorString(e));
  } else {
    e = ensure_dynamic_smem(tc::pw_tc_kernel<false>, optin_plain, smem, false);
    X3D_REQUIRE(e == cudaSuccess, X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: smem attribute: %s", cudaGetErrorString(e));
    e = launch_pdl(tc::pw_tc_kernel<false>, grid, dim3(tc::kThreadsPlain), smem, st, tmA, tmW, tmD, tmA2, p);
    X3D_REQUIRE(e == cudaSuccess, X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: launch: %s", cudaGetErrorString(e));
  }
  return check_launch("x3d_pw_tc_fwd");
}

#ifdef X3D_PW_TRACE
extern "C" int x3d_pw_trace_dump(long long* host_out) {      // [256][8] clock64 stamps of CTA 7, group 0
  return cudaMemcpyFromSymbol(host_out, x3d::tc::g_pw_trace, sizeof(long long) * 8 * 256) == cudaSuccess ? 0 : -3;
}
#endif
