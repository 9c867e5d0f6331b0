// Channelwise 3x3x3 convolution + folded BN (+ SE partial sums), TMA-staged, marching over T.
// Replaces Bottleneck.b + bn_b (+ the reduction of se_pool), reference model.py:309-312.
//
// One CTA = one clip x one chunk of CH channels x one spatial tile of Q rows x SW columns of the
// output.  The CTA walks the T input frames once: frame t's halo tile ((Q-1)*S+3) x ((SW-1)*S+3)
// x CH lands in a 3-slot shared-memory ring by ONE 5-D TMA box copy (out-of-range rows, columns
// and channels are zero-filled by the TMA unit = TF 'SAME' zero padding and channel padding, no
// predicates in the kernel).  A thread owns one channel pair (its 27 taps stay in registers) and
// one output row of the tile; every staged value is read from shared memory once per step and
// scattered into three accumulator sets (output frames t-1, t, t+1), i.e. 9 FMAs per loaded
// value, issued as packed FFMA2.  Frame t-1 is complete after step t; it is staged in shared
// memory and written back by one TMA box store (clipped at the tensor edge) while the next frames
// are already in flight.  All shared-memory offsets are compile-time constants; the kernel does
// no global address arithmetic at all.
#include <stdlib.h>
#include <string.h>

#include "tma_common.cuh"

namespace x3d {
namespace dwt {

using namespace ptx;

constexpr int kSlots = 3;

struct Params {
  const float* w;        // [27, Cs] BN-folded taps
  const float* bias;     // [Cs]
  float* partial;        // [N, tiles, Cs] or nullptr
  int T, Ho, Wo, Cs;
  int Q;                 // output rows per tile (= thread slots)
  int tiles_w, tiles;    // spatial tiles per frame
  int pad_h, pad_w;
  int slot_bytes;        // ring pitch: bytes of one staged frame tile rounded up to 128
  int box_bytes;         // exact bytes one TMA box copy delivers
  int stage_bytes;       // output staging tile Q*SW*CH elements, rounded up to 128
  int act;               // 1: swish applied to the output (blocks without SE, model.py:316)
};

// shared-memory reads by 32-bit shared-space address (keeps LDS, never generic LD)
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
    uint32_t u;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(a));
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
};

template <typename T> struct Pack;
template <> struct Pack<float> {
  static __device__ __forceinline__ void sts2(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
  }
};
template <> struct Pack<bf16> {
  static __device__ __forceinline__ void sts2(uint32_t a, float2 v) {
    __nv_bfloat162 h = __float22bfloat162_rn(v);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(*reinterpret_cast<uint32_t*>(&h)) : "memory");
  }
};

template <typename T, int S, int SW, int CH>
__global__ void __launch_bounds__(224, 2)
dw_tma_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
              const Params p) {
  constexpr int BW = (SW - 1) * S + 3;
  constexpr int ES = sizeof(T);
  constexpr int PS = CH * ES;              // bytes per staged pixel
  constexpr int RS = BW * PS;              // bytes per staged input row
  constexpr int C2 = CH / 2;

  extern __shared__ __align__(128) uint8_t dw_smem_raw[];
  const uint32_t raw_s = smem_u32(dw_smem_raw);
  const uint32_t smem_s = (raw_s + 127u) & ~127u;              // shared-space byte address
  uint8_t* smem = dw_smem_raw + (smem_s - raw_s);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kSlots]
  const uint32_t ring_s = smem_s + 128;                        // input ring  [kSlots][slot_bytes]
  const uint32_t stage_s = ring_s + kSlots * p.slot_bytes;     // output ring [kSlots][stage_bytes]
  float* s_red = reinterpret_cast<float*>(smem + 128 + kSlots * (p.slot_bytes + p.stage_bytes));
  // dt=2 taps live in shared memory ([9][C2] float2, read back conflict-free) to keep the kernel
  // at <= 128 registers, i.e. two 7-warp CTAs (4 warps per scheduler) resident per SM
  const uint32_t w2_s = stage_s + kSlots * p.stage_bytes + 9 * CH * 4 /*s_red: Q<=9 rows*/;

  const int tid = threadIdx.x;
  const int slot = tid / C2, cp = tid - slot * C2;
  const int n = blockIdx.z, c0 = blockIdx.y * CH;
  const int tile_h = blockIdx.x / p.tiles_w, tile_w = blockIdx.x - tile_h * p.tiles_w;
  const int ho0 = tile_h * p.Q, wo0 = tile_w * SW;
  const int c = c0 + 2 * cp;
  const bool in_slot = slot < p.Q;
  const bool on = in_slot && ho0 + slot < p.Ho && c < p.Cs;    // thread owns real outputs
  const int hi0 = ho0 * S - p.pad_h, wi0 = wo0 * S - p.pad_w;

  if (tid == 0) {
    prefetch_tmap(&tmIn);
    prefetch_tmap(&tmOut);
    for (int s = 0; s < kSlots; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();                                  // (launch_pdl) the input is the previous kernel's output
  if (tid == 0) {
    for (int f = 0; f < kSlots && f < p.T; ++f) {
      mbar_expect_tx(&full[f], static_cast<uint32_t>(p.box_bytes));
      tma_load_5d(ring_s + f * p.slot_bytes, &tmIn, c0, wi0, hi0, f, n, &full[f]);
    }
  }

  float2 wr[18];
  float2 bia = make_float2(0.f, 0.f);
  if (on) {
#pragma unroll
    for (int i = 0; i < 18; ++i) wr[i] = ld2(p.w + i * p.Cs + c);
    bia = ld2(p.bias + c);
  } else {
#pragma unroll
    for (int i = 0; i < 18; ++i) wr[i] = make_float2(0.f, 0.f);
  }
  const uint32_t w2_t = w2_s + static_cast<uint32_t>(cp) * 8;
  if (slot == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const float2 w = (c < p.Cs) ? ld2(p.w + (18 + i) * p.Cs + c) : make_float2(0.f, 0.f);
      Pack<float>::sts2(w2_t + i * (C2 * 8), w);
    }
  }
  __syncthreads();

  // Three accumulator sets rotate over output frames (set = frame % 3).  A set is (re)started by
  // the first tap of the dt=0 pass with the BN shift as addend; only frame 0 needs a preset.
  float2 acc[3][SW];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int j = 0; j < SW; ++j) acc[a][j] = bia;
  float2 ssum = make_float2(0.f, 0.f);

  const int aslot = in_slot ? slot : 0;       // spare threads of the last warp stay in range
  const uint32_t toff = static_cast<uint32_t>(aslot * S) * RS + static_cast<uint32_t>(cp) * 2 * ES;
  const uint32_t soff = static_cast<uint32_t>(aslot * SW) * PS + static_cast<uint32_t>(cp) * 2 * ES;
  int ncol = p.Wo - wo0;                      // valid columns of this strip
  if (ncol > SW) ncol = SW;

  // Stages output frame t_out (accumulator set A) in shared memory; thread 0 then TMA-stores it
  // (rows >= Ho, columns >= Wo and channels >= Cs are clipped by the TMA unit).
  auto stage_out = [&](float2 (&A)[SW], int t_out) {
    const uint32_t dst = stage_s + (t_out % kSlots) * p.stage_bytes + soff;
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
      for (int j = 0; j < SW; ++j) Pack<T>::sts2(dst + j * PS, A[j]);
    }
    if (on) {
#pragma unroll
      for (int j = 0; j < SW; ++j)
        if (j < ncol) ssum = __fadd2_rn(ssum, A[j]);
    }
    fence_proxy_async();
  };
  auto store_out = [&](int t_out) {            // thread 0, after the CTA-wide barrier
    tma_store_5d(&tmOut, stage_s + (t_out % kSlots) * p.stage_bytes, c0, wo0, ho0, t_out, n);
    tma_store_commit();
    tma_store_wait_read<1>();                  // the store issued one step earlier has drained
  };

  // One step: input frame t contributes tap dt=0 to output t+1 (set A0, restarted here), dt=1 to
  // output t (A1) and dt=2 to output t-1 (A2), which is complete afterwards.
  auto step = [&](int t, float2 (&A0)[SW], float2 (&A1)[SW], float2 (&A2)[SW]) {
    const int s = t % kSlots;
    mbar_wait(&full[s], static_cast<uint32_t>((t / kSlots) & 1));
    const uint32_t base = ring_s + s * p.slot_bytes + toff;
#pragma unroll
    for (int dh = 0; dh < 3; ++dh) {
      float2 w2[3];
#pragma unroll
      for (int dw = 0; dw < 3; ++dw) w2[dw] = Elem<float>::lds2(w2_t + (dh * 3 + dw) * (C2 * 8));
#pragma unroll
      for (int jj = 0; jj < BW; ++jj) {
        // one staged value feeds up to 3 output columns x 3 output frames, then dies
        const float2 x = Elem<T>::lds2(base + dh * RS + jj * PS);
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
    if (t >= 1) stage_out(A2, t - 1);
    __syncthreads();                       // slot s fully read; staged output frame complete
    if (tid == 0) {
      if (t + kSlots < p.T) {
        mbar_expect_tx(&full[s], static_cast<uint32_t>(p.box_bytes));
        tma_load_5d(ring_s + s * p.slot_bytes, &tmIn, c0, wi0, hi0, t + kSlots, n, &full[s]);
      }
      if (t >= 1) store_out(t - 1);
    }
  };

  // Whole triples in the loop (after three steps the accumulator sets are back in their original
  // roles, so they live in fixed registers: no rotation moves at the loop edge), T % 3 frames after it.
  int t = 0;
  for (; t + 3 <= p.T; t += 3) {
    step(t, acc[1], acc[0], acc[2]);
    step(t + 1, acc[2], acc[1], acc[0]);
    step(t + 2, acc[0], acc[2], acc[1]);
  }
  const int rem = p.T - t;
  if (rem >= 1) step(t, acc[1], acc[0], acc[2]);
  if (rem == 2) step(t + 1, acc[2], acc[1], acc[0]);
  // the last output frame never sees a dt=2 contribution (temporal zero padding)
  if (rem == 0) stage_out(acc[2], p.T - 1);
  else if (rem == 1) stage_out(acc[0], p.T - 1);
  else stage_out(acc[1], p.T - 1);
  // many waves of CTAs: let the next kernel's CTAs in only now, when this one is about to leave
  pdl_trigger();
  if (p.partial != nullptr && in_slot) {
    s_red[slot * CH + 2 * cp] = ssum.x;
    s_red[slot * CH + 2 * cp + 1] = ssum.y;
  }
  __syncthreads();
  if (tid == 0) {
    store_out(p.T - 1);
    tma_store_wait_read<0>();                // shared memory must outlive the bulk stores
  }
  if (p.partial != nullptr) {
    for (int ch = tid; ch < CH; ch += blockDim.x) {
      if (c0 + ch < p.Cs) {
        float a = 0.f;
        for (int k = 0; k < p.Q; ++k) a += s_red[k * CH + ch];
        p.partial[(static_cast<long>(n) * p.tiles + blockIdx.x) * p.Cs + c0 + ch] = a;
      }
    }
  }
}

// ------------------------------------------------------------------------------ host
struct Plan {
  int CH, SW, Q, threads, tiles_w, tiles_h, chunks, BH, BW, slot_bytes, box_bytes, stage_bytes;
  size_t smem;
};

// Picks the channel chunk (56/64/72), strip width (7/8) and rows per tile (<= 224 threads, and at
// most ~110 KB of shared memory so that two CTAs stay resident per SM) that minimise the padded
// output volume (the kernel is FMA-bound), with the staged input volume as a secondary cost.
static Plan make_plan(int H, int W, int Cs, int stride, int esize) {
  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  const int chs[3] = {56, 64, 72}, sws[2] = {8, 7};
  Plan best{};
  double best_cost = 1e300;
  for (int ci = 0; ci < 3; ++ci) {
    for (int si = 0; si < 2; ++si) {
      const int CH = chs[ci], SW = sws[si];
      int qmax = 224 / (CH / 2);
      if (qmax > 9) qmax = 9;
      if (qmax > Ho) qmax = Ho;
      for (int Q = qmax; Q >= 1; --Q) {
        Plan pl;
        pl.CH = CH; pl.SW = SW; pl.Q = Q;
        pl.chunks = (Cs + CH - 1) / CH;
        pl.BW = (SW - 1) * stride + 3;
        pl.BH = (Q - 1) * stride + 3;
        pl.box_bytes = pl.BH * pl.BW * CH * esize;
        pl.slot_bytes = (pl.box_bytes + 127) / 128 * 128;
        pl.stage_bytes = (Q * SW * CH * esize + 127) / 128 * 128;
        pl.smem = 128 /*align*/ + 128 /*barriers*/ +
                  (size_t)kSlots * (pl.slot_bytes + pl.stage_bytes) + (size_t)9 * CH * sizeof(float) /*s_red*/ +
                  (size_t)9 * CH * sizeof(float) /*dt=2 taps*/;
        if (pl.smem > 110 * 1024 && Q > 1) continue;
        pl.threads = (Q * (CH / 2) + 31) / 32 * 32;
        pl.tiles_w = (Wo + SW - 1) / SW;
        pl.tiles_h = (Ho + Q - 1) / Q;
        const double tiles = (double)pl.tiles_w * pl.tiles_h * pl.chunks;
        const double work = tiles * Q * SW * CH;
        const double staged = tiles * pl.BH * pl.BW * CH;
        const double cost = work + 0.1 * staged + 1e-3 * tiles;
        if (cost < best_cost) { best_cost = cost; best = pl; }
      }
    }
  }
  return best;
}

template <typename T, int S, int SW, int CH>
static int launch(const CUtensorMap& tm, const CUtensorMap& tmo, const Params& p, const Plan& pl, int N, cudaStream_t st) {
  auto kern = dw_tma_kernel<T, S, SW, CH>;
  static SmemOptIn optin;                              // one per template instance, per device inside
  const cudaError_t e = ensure_dynamic_smem(kern, optin, pl.smem);
  if (e != cudaSuccess) {
    set_error("x3d_dw3x3x3_fwd: smem attribute (%zu B): %s", pl.smem, cudaGetErrorString(e));
    return X3D_ERR_LAUNCH;
  }
  dim3 grid(pl.tiles_w * pl.tiles_h, pl.chunks, N);
  const cudaError_t le = launch_pdl(kern, grid, dim3(pl.threads), pl.smem, st, tm, tmo, p);
  if (le != cudaSuccess) {
    set_error("x3d_dw3x3x3_fwd: launch: %s", cudaGetErrorString(le));
    return X3D_ERR_LAUNCH;
  }
  return check_launch("x3d_dw3x3x3_fwd");
}

template <typename T, int S>
static int dispatch(const CUtensorMap& tm, const CUtensorMap& tmo, const Params& p, const Plan& pl, int N, cudaStream_t st) {
#define X3D_DWT(SWW, CHH) \
  if (pl.SW == SWW && pl.CH == CHH) return launch<T, S, SWW, CHH>(tm, tmo, p, pl, N, st)
  X3D_DWT(8, 56); X3D_DWT(8, 64); X3D_DWT(8, 72);
  X3D_DWT(7, 56); X3D_DWT(7, 64); X3D_DWT(7, 72);
#undef X3D_DWT
  set_error("x3d_dw3x3x3_fwd: no kernel for SW=%d CH=%d", pl.SW, pl.CH);
  return X3D_ERR_UNSUPPORTED;
}

}  // namespace dwt
}  // namespace x3d

using namespace x3d;

extern "C" int x3d_dw_partial_blocks(int T, int H, int W, int C, int stride, int dtype) {
  if (T <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || (stride != 1 && stride != 2)) return 0;
  if (dtype != X3D_F32 && dtype != X3D_BF16) return 0;
  const dwt::Plan pl = dwt::make_plan(H, W, C, stride, dtype == X3D_BF16 ? 2 : 4);
  return pl.tiles_w * pl.tiles_h;
}

extern "C" int x3d_dw3x3x3_fwd(const void* in, const float* w, const float* bias, void* out,
                               float* se_partial, int N, int T, int H, int W, int C, int stride,
                               int pad_h, int pad_w, int dtype, void* stream) {
  return x3d_dw3x3x3_act_fwd(in, w, bias, out, se_partial, N, T, H, W, C, stride, pad_h, pad_w, dtype, 0, stream);
}

extern "C" int x3d_dw3x3x3_act_fwd(const void* in, const float* w, const float* bias, void* out,
                                   float* se_partial, int N, int T, int H, int W, int C, int stride,
                                   int pad_h, int pad_w, int dtype, int act, void* stream) {
  X3D_REQUIRE(act == 0 || (act == 1 && se_partial == nullptr), X3D_ERR_INVALID_ARG,
              "x3d_dw3x3x3_act_fwd: act=%d (0, or 1 = swish without SE sums: with SE the scale comes first)", act);
  X3D_REQUIRE(in && w && bias && out, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_fwd: null pointer");
  X3D_REQUIRE(C > 0 && C % 8 == 0, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_fwd: C=%d not a multiple of 8", C);
  X3D_REQUIRE(stride == 1 || stride == 2, X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_fwd: stride %d", stride);
  X3D_REQUIRE(N > 0 && N <= 65535 && T > 0 && H > 0 && W > 0, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_fwd: bad extent");
  X3D_REQUIRE(pad_h >= 0 && pad_h <= 1 && pad_w >= 0 && pad_w <= 1, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_fwd: pad_before must be 0 or 1");
  X3D_REQUIRE(dtype == X3D_F32 || dtype == X3D_BF16, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_fwd: dtype %d", dtype);
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_fwd: input must be 16-byte aligned");
  X3D_REQUIRE(device_sm_count() > 0 && device_is_sm100(), X3D_ERR_NO_DEVICE, "x3d_dw3x3x3_fwd: needs an sm_100 device");
  EncodeTiledFn enc = tensor_map_encoder();
  X3D_REQUIRE(enc != nullptr, X3D_ERR_NO_DEVICE, "x3d_dw3x3x3_fwd: cuTensorMapEncodeTiled unavailable");
  const int es = dtype == X3D_BF16 ? 2 : 4;
  const dwt::Plan pl = dwt::make_plan(H, W, C, stride, es);
  X3D_REQUIRE(pl.chunks <= 65535, X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_fwd: too many channel chunks");
  X3D_REQUIRE((int)pl.smem <= device_max_smem(), X3D_ERR_UNSUPPORTED, "x3d_dw3x3x3_fwd: tile needs %zu B of shared memory", pl.smem);

  CUtensorMap tm;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es,
                           (cuuint64_t)T * H * W * C * es};
  cuuint32_t box[5] = {(cuuint32_t)pl.CH, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tm, dtype == X3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   5, const_cast<void*>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_dw3x3x3_fwd: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d,%d] box [%d,%d,%d]",
              (int)r, N, T, H, W, C, pl.CH, pl.BW, pl.BH);

  const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_dw3x3x3_fwd: output must be 16-byte aligned");
  CUtensorMap tmo;
  cuuint64_t odims[5] = {(cuuint64_t)C, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t ostrides[4] = {(cuuint64_t)C * es, (cuuint64_t)Wo * C * es, (cuuint64_t)Ho * Wo * C * es,
                            (cuuint64_t)T * Ho * Wo * C * es};
  cuuint32_t obox[5] = {(cuuint32_t)pl.CH, (cuuint32_t)pl.SW, (cuuint32_t)pl.Q, 1, 1};
  r = enc(&tmo, dtype == X3D_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
          5, out, odims, ostrides, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  X3D_REQUIRE(r == CUDA_SUCCESS, X3D_ERR_LAUNCH, "x3d_dw3x3x3_fwd: output tensor map failed (%d)", (int)r);

  dwt::Params p;
  p.w = w; p.bias = bias; p.partial = se_partial;
  p.T = T; p.Ho = Ho; p.Wo = Wo; p.Cs = C;
  p.Q = pl.Q; p.tiles_w = pl.tiles_w; p.tiles = pl.tiles_w * pl.tiles_h;
  p.pad_h = pad_h; p.pad_w = pad_w; p.slot_bytes = pl.slot_bytes; p.box_bytes = pl.box_bytes;
  p.stage_bytes = pl.stage_bytes;
  p.act = act;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == X3D_BF16)
    return stride == 1 ? dwt::dispatch<bf16, 1>(tm, tmo, p, pl, N, st) : dwt::dispatch<bf16, 2>(tm, tmo, p, pl, N, st);
  return stride == 1 ? dwt::dispatch<float, 1>(tm, tmo, p, pl, N, st) : dwt::dispatch<float, 2>(tm, tmo, p, pl, N, st);
}
