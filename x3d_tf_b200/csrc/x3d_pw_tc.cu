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
//   * 4 epilogue warps read TMEM with tcgen05.ld, add bias (+ residual), ReLU, pack bf16, store;
//   * optional prologue (projection conv of SE blocks, model.py:311-317): 4 transform warps apply
//     swish(se[clip,k] * a) in place on each landed stage before the MMA consumes it.
//
// Reference call sites replaced: Bottleneck.a/bn_a/relu (model.py:306-308), Bottleneck.c/bn_c +
// ResBlock add/relu (model.py:317-318,389-392), conv5 (model.py:117).
#include "tma_common.cuh"

namespace x3d {
namespace tc {

constexpr int kBlockM = 128, kBlockK = 64, kStageBytes = kBlockM * kBlockK * 2;
using namespace ptx;

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
  int KC;          // number of 64-wide K chunks
  int k16_last;    // number of K=16 MMAs in the last chunk (1..4)
  int stages;      // A ring depth
  int tmem_cols;   // power of two >= 2*NT
  int relu, swish;
};

constexpr int kThreadsPlain = 192, kThreadsPro = 320;

template <bool kPro>
__global__ void __launch_bounds__(kPro ? kThreadsPro : kThreadsPlain, 1)
pw_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * p.NT;
  const int w_chunk_bytes = p.NT * 128;

  uint8_t* sW = smem;
  uint8_t* sA = sW + p.KC * w_chunk_bytes;
  float* sBias = reinterpret_cast<float*>(sA + p.stages * kStageBytes);
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
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&xform[s], 128);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 128);
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

  const long num_tiles = (p.M + kBlockM - 1) / kBlockM;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_expect_tx(w_full, static_cast<uint32_t>(p.KC * w_chunk_bytes));
      for (int kc = 0; kc < p.KC; ++kc)
        tma_load_2d(sW + kc * w_chunk_bytes, &tmW, kc * kBlockK, n0, w_full);
      int s = 0;
      uint32_t ph = 0;
      for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int row0 = static_cast<int>(tile * kBlockM);
        for (int kc = 0; kc < p.KC; ++kc) {
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kStageBytes);
          tma_load_2d(sA + s * kStageBytes, &tmA, kc * kBlockK, row0, &full[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(p.NT);
    mbar_wait(w_full, 0);
    int s = 0;
    uint32_t ph = 0;
    long it = 0;
    for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
          const int nk = (kc == p.KC - 1) ? p.k16_last : 4;
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
  } else if (warp < 6) {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    long it = 0;
    for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int as = static_cast<int>(it & 1);
      const uint32_t aph = static_cast<uint32_t>((it >> 1) & 1);
      mbar_wait(&t_full[as], aph);
      tcgen05_after_sync();
      const long row = tile * kBlockM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * p.NT);
      bf16* drow = p.D + row * p.ldd + n0;
      const bf16* rrow = p.R ? p.R + row * p.ldr + n0 : nullptr;
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        if (n0 + c0 >= p.Nc) break;               // uniform across the CTA
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = c0 + h * 8;
          if (n0 + cc >= p.Nc || !row_ok) continue;
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(v[h * 8 + j]) + sBias[cc + j];
          if (rrow) {
            float r[8];
            ld8(rrow + cc, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] += r[j];
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
          }
          uint4 o;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(y[0], y[1]);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(y[2], y[3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(y[4], y[5]);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(y[6], y[7]);
          o.x = *reinterpret_cast<uint32_t*>(&t0);
          o.y = *reinterpret_cast<uint32_t*>(&t1);
          o.z = *reinterpret_cast<uint32_t*>(&t2);
          o.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(drow + cc) = o;
        }
      }
      tcgen05_before_sync();
      mbar_arrive(&t_empty[as]);
    }
  } else if (kPro) {
    // ------------------------------------------------------------------ prologue transform (warps 6..9)
    const int tt = threadIdx.x - 192;             // 0..127
    const int pchunk = tt & 7;                    // physical 16-byte chunk inside the 128-byte row
    int s = 0;
    uint32_t ph = 0;
    for (long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < p.KC; ++kc) {
        mbar_wait(&full[s], ph);
        uint8_t* st = sA + s * kStageBytes;
#pragma unroll 2
        for (int r0 = 0; r0 < kBlockM; r0 += 16) {
          const int r = r0 + (tt >> 3);
          const long row = tile * kBlockM + r;
          const int k = kc * kBlockK + ((pchunk ^ (r & 7)) << 3);
          if (row < p.M && k < p.Kc) {
            uint4* ptr = reinterpret_cast<uint4*>(st + r * 128 + pchunk * 16);
            uint4 u = *ptr;
            uint32_t w[4] = {u.x, u.y, u.z, u.w};
            float sc[8];
            if (p.se) {
              const float* sp = p.se + (row / p.rows_per_clip) * p.Kc + k;
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp));
              const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp) + 1);
              sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w;
              sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) sc[j] = 1.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a = __uint_as_float(w[j] << 16) * sc[2 * j];
              float b = __uint_as_float(w[j] & 0xffff0000u) * sc[2 * j + 1];
              if (p.swish) {
                float ta, tb;
                asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(0.5f * a));
                asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(0.5f * b));
                a = a * fmaf(0.5f, ta, 0.5f);
                b = b * fmaf(0.5f, tb, 0.5f);
              }
              __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
              w[j] = *reinterpret_cast<uint32_t*>(&t);
            }
            *ptr = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(&xform[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
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


}  // namespace tc
}  // namespace x3d

using namespace x3d;

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
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(a->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->Wp) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(a->D) & 15) == 0, X3D_ERR_INVALID_ARG, "x3d_pw_tc_fwd: pointers must be 16-byte aligned");
  X3D_REQUIRE(device_sm_count() > 0 && device_is_sm100(), X3D_ERR_NO_DEVICE, "x3d_pw_tc_fwd: needs an sm_100 device");
  // N tiling: fewest tiles with NT <= 256 whose resident weight slice leaves >= 3 A stages.
  const int kc_n = a->Kpad / 64;
  const int k16_total = (a->K + 15) / 16;
  const int KC = (k16_total + 3) / 4;                 // chunks that actually hold data
  const int k16_last = k16_total - (KC - 1) * 4;
  (void)kc_n;
  const int budget = device_max_smem() - 1024 /*align*/ - 1024 /*bias*/ - 512 /*barriers*/;
  int n_tiles = (a->Npad + 255) / 256;
  int NT = 0, stages = 0;
  for (;; ++n_tiles) {
    NT = (((a->Npad + n_tiles - 1) / n_tiles) + 15) / 16 * 16;
    const int wbytes = KC * NT * 128;
    stages = (budget - wbytes) / tc::kStageBytes;
    if (stages >= 3 || NT <= 16) break;
  }
  X3D_REQUIRE(stages >= 2, X3D_ERR_UNSUPPORTED, "x3d_pw_tc_fwd: K=%d too large for shared memory", a->K);
  if (stages > 8) stages = 8;
  X3D_REQUIRE((long)n_tiles * NT <= a->Npad + 15 || n_tiles * NT <= ((a->Npad + 15) / 16) * 16 + 16 * n_tiles,
              X3D_ERR_UNSUPPORTED, "x3d_pw_tc_fwd: tiling error");
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

  tc::Params p;
  p.bias = a->bias; p.R = static_cast<const bf16*>(a->R); p.se = a->se; p.D = static_cast<bf16*>(a->D);
  p.M = a->M; p.rows_per_clip = a->rows_per_clip; p.Kc = a->K; p.Nc = a->Nc; p.ldr = a->ldr; p.ldd = a->ldd;
  p.NT = NT; p.KC = KC; p.k16_last = k16_last; p.stages = stages; p.tmem_cols = tmem_cols;
  p.relu = a->relu; p.swish = a->swish;

  const size_t smem = 1024 + (size_t)KC * NT * 128 + (size_t)stages * tc::kStageBytes + 1024 + 512;
  const long num_tiles = (a->M + tc::kBlockM - 1) / tc::kBlockM;
  int gx = device_sm_count() / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > num_tiles) gx = (int)num_tiles;
  dim3 grid(gx, n_tiles);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool pro = (a->se != nullptr) || a->swish;
  cudaError_t e;
  if (pro) {
    e = cudaFuncSetAttribute(tc::pw_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    X3D_REQUIRE(e == cudaSuccess, X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: smem attribute: %s", cudaGetErrorString(e));
    tc::pw_tc_kernel<true><<<grid, tc::kThreadsPro, smem, st>>>(tmA, tmW, p);
  } else {
    e = cudaFuncSetAttribute(tc::pw_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    X3D_REQUIRE(e == cudaSuccess, X3D_ERR_LAUNCH, "x3d_pw_tc_fwd: smem attribute: %s", cudaGetErrorString(e));
    tc::pw_tc_kernel<false><<<grid, tc::kThreadsPlain, smem, st>>>(tmA, tmW, p);
  }
  return check_launch("x3d_pw_tc_fwd");
}
