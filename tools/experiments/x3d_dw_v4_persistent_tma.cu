// Channelwise 3x3x3 convolution + folded BN (+ SE partial sums), generation 4.
// Replaces Bottleneck.b + bn_b (+ the reduction of se_pool), reference model.py:309-312.
//
// Same data movement as x3d_dw_tma.cu (one 5-D TMA box per input frame into a 3-slot ring, OOB
// zero fill = TF 'SAME' padding, march over T with three rotating accumulator sets, finished
// frame staged in shared memory and written by one TMA box store), but built around the fact that
// the kernel is bound by the fp32 FMA pipe AND by issue slots at the same time (27 MACs per 3.2
// bytes; a packed FFMA2 holds the pipe for two cycles, so every non-FMA instruction beyond one
// per FFMA2 costs time):
//   * a thread owns one channel pair and a 2 x 4 patch of outputs: its input window is 4 x 6
//     (stride 1) or 5 x 9 (stride 2) values per frame instead of 3 x 10 / 3 x 17 for a 1 x 8
//     row, i.e. 0.33 (0.63) shared-memory loads + unpacks per FFMA2 instead of 0.42 (0.71);
//   * all 27 taps of the pair live in registers (no per-step tap reloads);
//   * no CTA-wide barrier in the frame loop: warps hand slots back through an mbarrier
//     (`done`), warp 0 alone waits for it and re-arms the TMA load / issues the TMA store, so a
//     fast warp runs up to two frames ahead of a slow one instead of idling at bar.sync;
//   * the SE reduction is a template switch (half of the blocks have no SE).
#include <stdlib.h>
#include <string.h>

#include "tma_common.cuh"

namespace x3d {
namespace dw4 {

using namespace ptx;

constexpr int kMaxIn = 8;     // input ring depth (frames in flight) is a launch parameter <= 8;
                              // the output staging ring is 3 deeper (slot-reuse argument in the loop)
static int ring_depth() {
  static const int d = [] {
    const char* e = getenv("X3D_DW4_KIN");
    const int v = e ? atoi(e) : 4;
    return v < 2 ? 2 : (v > kMaxIn ? kMaxIn : v);
  }();
  return d;
}
constexpr int RH = 2, RW = 4; // outputs per thread
// Register allocation is per 4 warps, so a 7-warp CTA pays for 8.  Two shapes are built:
//   kBig = false: up to 256 threads at <= 128 registers, two CTAs per SM;
//   kBig = true : up to 384 threads at <= 168 registers, one CTA per SM.
constexpr int kThreadsSmall = 256, kThreadsBig = 384;

struct Params {
  const float* w;        // [27, Cs] BN-folded taps
  const float* bias;     // [Cs]
  float* partial;        // [N, tiles, Cs] or nullptr
  int T, Ho, Wo, Cs;
  int ncg, nslots;       // column groups per tile, active thread slots (row groups x column groups)
  int QH, QW;            // output rows / columns per tile
  int tiles_w, tiles;
  int units;             // N * tiles: (clip, spatial tile) pairs per channel chunk
  int pad_h, pad_w;
  int row_bytes;         // pitch of one staged input row
  int slot_bytes, box_bytes, stage_bytes;
  int nwarps;
  int kin, kout;         // ring depths
  int skip;              // debug: 1 = no FMAs, 2 = no FMAs and no staging stores (pipeline only)
};

template <typename T> struct Io;
template <> struct Io<float> {
  static __device__ __forceinline__ float2 lds2(uint32_t a) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
    return r;
  }
  static __device__ __forceinline__ void sts2(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
  }
};
template <> struct Io<bf16> {
  static __device__ __forceinline__ float2 lds2(uint32_t a) {
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(a));
    uint32_t lo, hi;                          // byte permutes stay on the ALU pipe; a shift may be
    asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(lo) : "r"(u));   // turned into IMAD (FMA pipe)
    asm("prmt.b32 %0, %1, 0, 0x3244;" : "=r"(hi) : "r"(u));
    return make_float2(__uint_as_float(lo), __uint_as_float(hi));
  }
  static __device__ __forceinline__ void sts2(uint32_t a, float2 v) {
    __nv_bfloat162 h = __float22bfloat162_rn(v);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(*reinterpret_cast<uint32_t*>(&h)) : "memory");
  }
};

template <typename T, int S, int CH, bool SE, bool kBig>
__global__ void __launch_bounds__(kBig ? kThreadsBig : kThreadsSmall, kBig ? 1 : 2)
dw4_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
           const Params p) {
  constexpr int ES = sizeof(T);
  constexpr int PS = CH * ES;                 // bytes per staged pixel
  constexpr int C2 = CH / 2;
  constexpr int WR = (RH - 1) * S + 3;        // input window of one thread
  constexpr int WC = (RW - 1) * S + 3;

  const int kIn = p.kin, kOut = p.kout;
  extern __shared__ __align__(128) uint8_t dw4_smem_raw[];
  const uint32_t raw_s = smem_u32(dw4_smem_raw);
  const uint32_t smem_s = (raw_s + 127u) & ~127u;
  uint8_t* smem = dw4_smem_raw + (smem_s - raw_s);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kIn]  TMA -> warps
  uint64_t* done = full + kMaxIn;                                 // [kIn]  warps -> warp 0
  const uint32_t ring_s = smem_s + 128;
  const uint32_t stage_s = ring_s + kIn * p.slot_bytes;
  // per-unit SE sums of every thread slot, kIn deep (a warp is never more than kIn steps ahead
  // of warp 0, which drains them)
  float* s_red = reinterpret_cast<float*>(smem + 128 + kIn * p.slot_bytes + kOut * p.stage_bytes);

  const int tid = threadIdx.x, lane = tid & 31;
  int slot = tid / C2;
  const int cp = tid - slot * C2;
  const bool active = slot < p.nslots;
  if (!active) slot = 0;                      // spare lanes of the last warp stay in range
  const int rg = slot / p.ncg, cg = slot - rg * p.ncg;
  const int c0 = blockIdx.y * CH;
  const int c = c0 + 2 * cp;
  const bool on = active && c < p.Cs;

  // ---- work list: units (clip, spatial tile) u0, u0 + gridDim.x, ... of this channel chunk.
  // The CTA is persistent: the frame pipeline (TMA ring, staging ring, barriers) runs straight
  // through unit boundaries, only the accumulators restart.
  const int u0 = blockIdx.x, ustride = gridDim.x;
  const int nunits = (p.units - u0 + ustride - 1) / ustride;      // >= 1 (grid is clamped to units)
  const int G = nunits * p.T;                                     // frames this CTA consumes

  if (tid == 0) {
    prefetch_tmap(&tmIn);
    prefetch_tmap(&tmOut);
    for (int s = 0; s < kIn; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], static_cast<uint32_t>(p.nwarps));
    }
    fence_barrier_init();
  }
  __syncthreads();

  // loader state (thread 0 only): position of the next frame to request
  int ld_u = u0, ld_t = 0, ld_g = 0;
  auto issue_load = [&]() {                    // thread 0
    const int n = ld_u / p.tiles, tile = ld_u - n * p.tiles;
    const int th = tile / p.tiles_w, tw = tile - th * p.tiles_w;
    const int s = ld_g % kIn;
    mbar_expect_tx(&full[s], static_cast<uint32_t>(p.box_bytes));
    tma_load_5d(ring_s + s * p.slot_bytes, &tmIn, c0, tw * p.QW * S - p.pad_w, th * p.QH * S - p.pad_h,
                ld_t, n, &full[s]);
    ++ld_g;
    if (++ld_t == p.T) { ld_t = 0; ld_u += ustride; }
  };
  if (tid == 0) {
    for (int f = 0; f < kIn && f < G; ++f) issue_load();
  }

  float2 w[27];
  float2 bia = make_float2(0.f, 0.f);
  if (on) {
#pragma unroll
    for (int i = 0; i < 27; ++i) w[i] = ld2(p.w + i * p.Cs + c);
    bia = ld2(p.bias + c);
  } else {
#pragma unroll
    for (int i = 0; i < 27; ++i) w[i] = make_float2(0.f, 0.f);
  }

  const uint32_t toff = static_cast<uint32_t>(rg * RH * S) * p.row_bytes +
                        static_cast<uint32_t>(cg * RW * S) * PS + static_cast<uint32_t>(cp) * 2 * ES;
  const uint32_t out_pitch = static_cast<uint32_t>(p.QW) * PS;
  const uint32_t soff = static_cast<uint32_t>(rg * RH) * out_pitch +
                        static_cast<uint32_t>(cg * RW) * PS + static_cast<uint32_t>(cp) * 2 * ES;

  // pipeline positions, advanced identically by every thread
  int in_s = 0;  uint32_t in_ph = 0;           // input slot / parity of the frame being consumed
  int st_s = 0;                                // staging slot of the next finished output frame
  int sq_s = 0;                                // thread 0: staging slot of the next frame to store
  float2 ssum = make_float2(0.f, 0.f), ssum1 = make_float2(0.f, 0.f);
  uint32_t vmask = 0;                          // bit r*RW+j: output (r, j) of the patch is real

  float2 acc[3][RH][RW];

  auto stage_out = [&](float2 (&A)[RH][RW]) {
    const uint32_t dst = stage_s + st_s * p.stage_bytes + soff;
    if (++st_s == kOut) st_s = 0;
    if (active) {
#pragma unroll
      for (int r = 0; r < RH; ++r)
#pragma unroll
        for (int j = 0; j < RW; ++j) Io<T>::sts2(dst + r * out_pitch + j * PS, A[r][j]);
    }
    if (SE) {
#pragma unroll
      for (int r = 0; r < RH; ++r)
#pragma unroll
        for (int j = 0; j < RW; ++j)
          if (vmask & (1u << (r * RW + j))) {
            if (j & 1) ssum1 = __fadd2_rn(ssum1, A[r][j]);
            else ssum = __fadd2_rn(ssum, A[r][j]);
          }
    }
  };

  for (int ui = 0; ui < nunits; ++ui) {
    const int u = u0 + ui * ustride;
    const int n = u / p.tiles, tile = u - n * p.tiles;
    const int tile_h = tile / p.tiles_w, tile_w = tile - tile_h * p.tiles_w;
    const int ho0 = tile_h * p.QH, wo0 = tile_w * p.QW;
    if (SE) {
      vmask = 0;
      if (on) {
        const int nrv = min(RH, max(0, p.Ho - (ho0 + rg * RH)));
        const int ncv = min(RW, max(0, p.Wo - (wo0 + cg * RW)));
        for (int r = 0; r < nrv; ++r) vmask |= ((1u << ncv) - 1u) << (r * RW);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int r = 0; r < RH; ++r)
#pragma unroll
        for (int j = 0; j < RW; ++j) acc[a][r][j] = bia;

    // One step: input frame t contributes tap dt=0 to output t+1 (set A0, restarted here), dt=1
    // to output t (A1) and dt=2 to output t-1 (A2), which is complete afterwards; after the last
    // frame of the unit A1 is complete as well (temporal zero padding).
    //
    // Slot reuse without a CTA barrier.  Input slot: re-armed by thread 0 only after every warp
    // arrived on done[] for the step that read it.  Staging slot (kOut = kIn + 3): a warp that is
    // past the full[] wait of frame g knows thread 0 reached the epilogue of step g - kIn, hence
    // finished the epilogue of step g - kIn - 1, whose `wait_group.read 1` drained every store
    // group up to step g - kIn - 2, i.e. every output frame up to sequence number g - kIn - 2; at
    // the end of step g it writes sequence numbers <= g, whose slots last held <= g - kOut.
    auto step = [&](int t, float2 (&A0)[RH][RW], float2 (&A1)[RH][RW], float2 (&A2)[RH][RW]) {
      mbar_wait_lean(&full[in_s], in_ph);
      uint32_t rowa = ring_s + in_s * p.slot_bytes + toff;
      if (p.skip == 0)
#pragma unroll
      for (int r = 0; r < WR; ++r) {
#pragma unroll
        for (int ci = 0; ci < WC; ++ci) {
          const float2 x = Io<T>::lds2(rowa + ci * PS);
#pragma unroll
          for (int ro = 0; ro < RH; ++ro) {
            const int dh = r - ro * S;
            if (dh < 0 || dh > 2) continue;
#pragma unroll
            for (int co = 0; co < RW; ++co) {
              const int dw = ci - co * S;
              if (dw < 0 || dw > 2) continue;
              A0[ro][co] = fma2(x, w[(0 * 3 + dh) * 3 + dw], (dh == 0 && dw == 0) ? bia : A0[ro][co]);
              A1[ro][co] = fma2(x, w[(1 * 3 + dh) * 3 + dw], A1[ro][co]);
              A2[ro][co] = fma2(x, w[(2 * 3 + dh) * 3 + dw], A2[ro][co]);
            }
          }
        }
        rowa += p.row_bytes;
      }
      const bool last = t == p.T - 1;
      if (t >= 1 && p.skip < 2) stage_out(A2);
      if (last) {
        if (p.skip < 2) stage_out(A1);
        if (SE && active) {
          ssum = __fadd2_rn(ssum, ssum1);
          float* r = s_red + ((ui % kIn) * p.nslots + slot) * CH + 2 * cp;
          r[0] = ssum.x;
          r[1] = ssum.y;
          ssum = make_float2(0.f, 0.f);
          ssum1 = make_float2(0.f, 0.f);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[in_s]);
      if (tid < 32) {
        if (tid == 0) {
          mbar_wait_lean(&done[in_s], in_ph);  // every warp has read the slot and staged its outputs
          if (ld_g < G) issue_load();          // first: the refill is what the other warps wait for
          if (t >= 1 && p.skip < 2) {
            tma_store_5d(&tmOut, stage_s + sq_s * p.stage_bytes, c0, wo0, ho0, t - 1, n);
            if (++sq_s == kOut) sq_s = 0;
          }
          if (last && p.skip < 2) {
            tma_store_5d(&tmOut, stage_s + sq_s * p.stage_bytes, c0, wo0, ho0, t, n);
            if (++sq_s == kOut) sq_s = 0;
          }
          tma_store_commit();                          // one bulk group per step (may be empty)
          tma_store_wait_read<1>();                    // every group but this step's has drained
        }
        if (SE && last) {
          __syncwarp();
          const float* r = s_red + (ui % kIn) * p.nslots * CH;
          for (int ch = lane; ch < CH; ch += 32) {
            if (c0 + ch < p.Cs) {
              float a = 0.f;
              for (int k = 0; k < p.nslots; ++k) a += r[k * CH + ch];
              p.partial[(static_cast<long>(n) * p.tiles + tile) * p.Cs + c0 + ch] = a;
            }
          }
        }
      }
      __syncwarp();
      if (++in_s == kIn) { in_s = 0; in_ph ^= 1u; }
    };

    for (int t = 0; t < p.T; t += 3) {
      step(t, acc[1], acc[0], acc[2]);
      if (t + 1 < p.T) step(t + 1, acc[2], acc[1], acc[0]);
      if (t + 2 < p.T) step(t + 2, acc[0], acc[2], acc[1]);
    }
  }
  if (tid == 0) tma_store_wait_read<0>();      // shared memory must outlive the bulk stores
}

// ------------------------------------------------------------------------------ host
struct Plan {
  int CH, nrg, ncg, QH, QW, threads, tiles_w, tiles_h, chunks, BH, BW;
  int slot_bytes, box_bytes, stage_bytes;
  size_t smem;
};

static size_t smem_bytes(const Plan& pl) {
  const int kIn = ring_depth(), kOut = kIn + 3;
  return 128 /*align*/ + 128 /*barriers*/ + (size_t)kIn * pl.slot_bytes + (size_t)kOut * pl.stage_bytes +
         (size_t)kIn * pl.nrg * pl.ncg * pl.CH * sizeof(float) /*s_red*/;
}

// Picks channel chunk (56/64/72) and the tile (row groups x column groups of 2x4 patches).
// Cost = issued FMA volume (padded outputs incl. idle lanes) / an occupancy factor, plus a small
// weight on the staged input volume (halo re-reads come from L2).
static bool use_big() {
  static const bool big = [] {
    const char* e = getenv("X3D_DW4_VAR");
    return !(e != nullptr && strcmp(e, "A") == 0);
  }();
  return big;
}

static Plan make_plan(int H, int W, int Cs, int stride, int esize) {
  const int max_threads = use_big() ? kThreadsBig : kThreadsSmall;
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const int chs[3] = {56, 64, 72};
  Plan best{};
  double best_cost = 1e300;
  for (int ci = 0; ci < 3; ++ci) {
    const int CH = chs[ci], C2 = CH / 2;
    for (int nrg = 1; nrg <= 8; ++nrg) {
      for (int ncg = 1; ncg <= 8; ++ncg) {
        Plan pl;
        pl.CH = CH; pl.nrg = nrg; pl.ncg = ncg;
        pl.QH = nrg * RH; pl.QW = ncg * RW;
        if (pl.QH > Ho + RH - 1 && nrg > 1) continue;      // tile taller than the frame
        if (pl.QW > Wo + RW - 1 && ncg > 1) continue;
        const int lanes = nrg * ncg * C2;
        pl.threads = (lanes + 31) / 32 * 32;
        if (pl.threads > max_threads) continue;
        pl.chunks = (Cs + CH - 1) / CH;
        pl.BW = (pl.QW - 1) * stride + 3;
        pl.BH = (pl.QH - 1) * stride + 3;
        if (pl.BW > 256 || pl.BH > 256) continue;
        pl.box_bytes = pl.BH * pl.BW * CH * esize;
        pl.slot_bytes = (pl.box_bytes + 127) / 128 * 128;
        pl.stage_bytes = (pl.QH * pl.QW * CH * esize + 127) / 128 * 128;
        pl.smem = smem_bytes(pl);
        if (pl.smem > 110 * 1024) continue;
        pl.tiles_w = (Wo + pl.QW - 1) / pl.QW;
        pl.tiles_h = (Ho + pl.QH - 1) / pl.QH;
        // resident warps per SM: registers (<= 144/thread -> 14 warps), shared memory, 32 CTAs
        int ctas = (int)((227 * 1024) / (pl.smem + 1024));
        const int alloc = (pl.threads + 127) / 128 * 128;            // warps are allocated in fours
        const int by_regs = use_big() ? 384 / alloc : 512 / alloc;
        if (by_regs < ctas) ctas = by_regs;
        if (ctas < 1) continue;
        const int warps = ctas * pl.threads / 32;
        const double occ = warps >= 12 ? 1.0 : (warps >= 8 ? 0.85 : 0.6);
        const double tiles = (double)pl.tiles_w * pl.tiles_h * pl.chunks;
        const double work = tiles * pl.threads * RH * RW * 2;      // issued output channels
        const double staged = tiles * pl.BH * pl.BW * CH;
        const double cost = (work + 0.15 * staged) / occ + 1e-3 * tiles;
        if (cost < best_cost) { best_cost = cost; best = pl; }
      }
    }
  }
  return best;
}

template <typename T, int S, int CH, bool SE, bool kBig>
static int launch(const CUtensorMap& tm, const CUtensorMap& tmo, const Params& p, const Plan& pl, int N, cudaStream_t st) {
  auto kern = dw4_kernel<T, S, CH, SE, kBig>;
  static size_t configured = 0;
  if (pl.smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) {
      set_error("x3d_dw3x3x3_fwd: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e));
      return X3D_ERR_LAUNCH;
    }
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = pl.smem;
  }
  // persistent grid: as many CTAs per channel chunk as fit on the chip at once, evened out so
  // that every CTA walks the same number of (clip, tile) units (+-1)
  static int occ_threads = 0, occ_ctas = 0;
  static size_t occ_smem = 0;
  if (occ_threads != pl.threads || occ_smem != pl.smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, pl.threads, pl.smem) != cudaSuccess || nb < 1) nb = 1;
    occ_threads = pl.threads; occ_smem = pl.smem; occ_ctas = nb;
  }
  int per_chunk = occ_ctas * device_sm_count() / pl.chunks;
  if (per_chunk < 1) per_chunk = 1;
  { const char* e = getenv("X3D_DW4_WAVES"); if (e) per_chunk = (int)(per_chunk * atof(e)); }   // debug
  const int waves = (p.units + per_chunk - 1) / per_chunk;
  const int gx = (p.units + waves - 1) / waves;
  dim3 grid(gx, pl.chunks, 1);
  (void)N;
  kern<<<grid, pl.threads, pl.smem, st>>>(tm, tmo, p);
  return check_launch("x3d_dw3x3x3_fwd");
}

template <typename T, int S>
static int dispatch(const CUtensorMap& tm, const CUtensorMap& tmo, const Params& p, const Plan& pl, int N, cudaStream_t st) {
  const bool se = p.partial != nullptr, big = use_big();
#define X3D_DW4(CHH)                                                                          \
  if (pl.CH == CHH)                                                                           \
    return se ? (big ? launch<T, S, CHH, true, true>(tm, tmo, p, pl, N, st)                   \
                     : launch<T, S, CHH, true, false>(tm, tmo, p, pl, N, st))                 \
              : (big ? launch<T, S, CHH, false, true>(tm, tmo, p, pl, N, st)                  \
                     : launch<T, S, CHH, false, false>(tm, tmo, p, pl, N, st))
  X3D_DW4(56); X3D_DW4(64); X3D_DW4(72);
#undef X3D_DW4
  set_error("x3d_dw3x3x3_fwd: no kernel for CH=%d", pl.CH);
  return X3D_ERR_UNSUPPORTED;
}

// Tile plan of a shape, for tests / documentation: {CH, row groups, column groups, threads,
// shared-memory bytes, spatial tiles, channel chunks, CTAs per SM assumed}.
void plan_debug(int H, int W, int C, int stride, int dtype, int* out) {
  const Plan pl = make_plan(H, W, C, stride, dtype == X3D_BF16 ? 2 : 4);
  out[0] = pl.CH; out[1] = pl.nrg; out[2] = pl.ncg; out[3] = pl.threads; out[4] = (int)pl.smem;
  out[5] = pl.tiles_w * pl.tiles_h; out[6] = pl.chunks; out[7] = use_big() ? 1 : 2;
}

int partial_blocks(int T, int H, int W, int C, int stride, int dtype) {
  const Plan pl = make_plan(H, W, C, stride, dtype == X3D_BF16 ? 2 : 4);
  return pl.tiles_w * pl.tiles_h;
}

int fwd(const void* in, const float* w, const float* bias, void* out, float* se_partial, int N, int T,
        int H, int W, int C, int stride, int pad_h, int pad_w, int dtype, void* stream) {
  EncodeTiledFn enc = tensor_map_encoder();
  X3D_REQUIRE(enc != nullptr, X3D_ERR_NO_DEVICE, "x3d_dw3x3x3_fwd: cuTensorMapEncodeTiled unavailable");
  const int es = dtype == X3D_BF16 ? 2 : 4;
  const Plan pl = make_plan(H, W, C, stride, es);
  X3D_REQUIRE(pl.threads > 0, X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_fwd: no tile plan for [%d,%d,%d] stride %d", H, W, C, stride);
  X3D_REQUIRE(pl.chunks <= 65535, X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_fwd: too many channel chunks");
  X3D_REQUIRE((int)pl.smem <= device_max_smem(), X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_fwd: tile needs %zu B of shared memory", pl.smem);
  const CUtensorMapDataType dt = dtype == X3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;

  CUtensorMap tm;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es,
                           (cuuint64_t)T * H * W * C * es};
  cuuint32_t box[5] = {(cuuint32_t)pl.CH, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tm, dt, 5, const_cast<void*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_dw3x3x3_fwd: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d,%d] box [%d,%d,%d]",
              (int)r, N, T, H, W, C, pl.CH, pl.BW, pl.BH);

  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  CUtensorMap tmo;
  cuuint64_t odims[5] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t ostrides[4] = {(cuuint64_t)C * es, (cuuint64_t)Wo * C * es, (cuuint64_t)Ho * Wo * C * es,
                            (cuuint64_t)T * Ho * Wo * C * es};
  cuuint32_t obox[5] = {(cuuint32_t)pl.CH, (cuuint32_t)pl.QW, (cuuint32_t)pl.QH, 1, 1};
  r = enc(&tmo, dt, 5, out, odims, ostrides, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_dw3x3x3_fwd: output tensor map failed (%d)", (int)r);

  Params p;
  p.w = w; p.bias = bias; p.partial = se_partial;
  p.T = T; p.Ho = Ho; p.Wo = Wo; p.Cs = C;
  p.ncg = pl.ncg; p.nslots = pl.nrg * pl.ncg; p.QH = pl.QH; p.QW = pl.QW;
  p.tiles_w = pl.tiles_w; p.tiles = pl.tiles_w * pl.tiles_h;
  p.units = N * p.tiles;
  p.pad_h = pad_h; p.pad_w = pad_w;
  p.row_bytes = pl.BW * pl.CH * es;
  p.slot_bytes = pl.slot_bytes; p.box_bytes = pl.box_bytes; p.stage_bytes = pl.stage_bytes;
  p.nwarps = pl.threads / 32;
  p.kin = ring_depth(); p.kout = p.kin + 3;
  { const char* e = getenv("X3D_DW4_SKIP"); p.skip = e ? atoi(e) : 0; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == X3D_BF16)
    return stride == 1 ? dispatch<bf16, 1>(tm, tmo, p, pl, N, st) : dispatch<bf16, 2>(tm, tmo, p, pl, N, st);
  return stride == 1 ? dispatch<float, 1>(tm, tmo, p, pl, N, st) : dispatch<float, 2>(tm, tmo, p, pl, N, st);
}

}  // namespace dw4
}  // namespace x3d
