// The two steps either side of the forward path that run on the device:
//   * input stage  -- utils.normalize (reference utils.py:42-72) on decoded uint8 frames;
//   * eval metrics -- the loss and metrics eval.py:62-70 compiles into the model
//     (SparseCategoricalCrossentropy on probabilities, SparseCategoricalAccuracy,
//     SparseTopKCategoricalAccuracy(k)).
#include "common.cuh"

namespace x3d {
namespace io {

struct Norm {
  float mean[3], std[3], nv;
};

// Same fp32 operation order as the reference: x / norm_value, then (x - mean) / std.
__device__ __forceinline__ float normalize1(float u, float nv, float mean, float std) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(u, nv), mean), std);
}

// A pixel value can only be one of 256 per channel, so every CTA first builds the 3 x 256 table of
// normalised values (exact IEEE fp32 in the reference's order, rounded once for bf16 output) in
// shared memory and the streaming loop is a byte extract + table read per element: 12 bytes in
// (three 32-bit loads = 4 pixels), 12 values out (16-byte stores) per thread and iteration.
template <typename TO>
__global__ void __launch_bounds__(256)
normalize_u8_kernel(const uint8_t* __restrict__ in, TO* __restrict__ out, int64_t pixels, const Norm nrm) {
  __shared__ TO lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int ch = i >> 8;
    const float r = normalize1(static_cast<float>(i & 255), nrm.nv, nrm.mean[ch], nrm.std[ch]);
    if (sizeof(TO) == 4) reinterpret_cast<float*>(&lut[0][0])[i] = r;
    else reinterpret_cast<bf16*>(&lut[0][0])[i] = __float2bfloat16_rn(r);
  }
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t groups = pixels / 4;
  for (int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; g < groups; g += stride) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(in + g * 12);
    const uint32_t w[3] = {__ldg(src), __ldg(src + 1), __ldg(src + 2)};
    TO v[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) v[i] = lut[i % 3][(w[i >> 2] >> (8 * (i & 3))) & 0xffu];
    if (sizeof(TO) == 4) {
      float4* dst = reinterpret_cast<float4*>(out + g * 12);
      const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
      for (int i = 0; i < 3; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    } else {
      // 24 bytes: 8-byte aligned for every g, 16-byte aligned for even g
      uint2* dst = reinterpret_cast<uint2*>(out + g * 12);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
#pragma unroll
      for (int i = 0; i < 3; ++i) dst[i] = make_uint2(u[2 * i], u[2 * i + 1]);
    }
  }
  // tail (pixels % 4) by the first threads of block 0
  if (blockIdx.x == 0) {
    const int64_t done = groups * 4;
    for (int64_t e = done * 3 + threadIdx.x; e < pixels * 3; e += blockDim.x) out[e] = lut[e % 3][in[e]];
  }
}

// One warp per video.  Keras semantics (TF 2.4 backend.sparse_categorical_crossentropy with
// from_logits=False on a tensor that is not a Softmax op output -- eval-mode X3D.call returns a
// mean over views): p = clip(p, 1e-7, 1 - 1e-7); loss = -(log p[label] - log sum_j p[j]).
// top-1: argmax (first maximal index) == label.  top-k: tf.math.in_top_k, i.e. fewer than k
// classes have a strictly larger probability.
__global__ void __launch_bounds__(128)
eval_metrics_kernel(const float* __restrict__ probs, const int32_t* __restrict__ labels,
                    double* __restrict__ acc, int V, int ncls, int k) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= V) return;
  const float* p = probs + static_cast<int64_t>(warp) * ncls;
  const int label = labels[warp];
  if (label < 0 || label >= ncls) return;                   // counted as a miss with zero loss
  const float pl = p[label];
  int greater = 0, before = 0;
  float sum = 0.f;
  for (int j = lane; j < ncls; j += 32) {
    const float v = p[j];
    greater += v > pl;
    before += (v == pl && j < label);
    sum += fminf(fmaxf(v, 1e-7f), 1.f - 1e-7f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    greater += __shfl_xor_sync(0xffffffffu, greater, o);
    before += __shfl_xor_sync(0xffffffffu, before, o);
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
  }
  if (lane == 0) {
    const float plc = fminf(fmaxf(pl, 1e-7f), 1.f - 1e-7f);
    atomicAdd(acc + 0, static_cast<double>(logf(sum) - logf(plc)));
    atomicAdd(acc + 1, (greater == 0 && before == 0) ? 1.0 : 0.0);
    atomicAdd(acc + 2, greater < k ? 1.0 : 0.0);
    atomicAdd(acc + 3, 1.0);
  }
}

// Evaluation clips of one decoded video: temporal views (transforms.py:48-65) and uniform spatial
// crops (transforms.py:149-190, 216-222) as ONE gather on the device.
//   out[crop][view][t][y][x][c] = in[((view*T + t) * rate) mod F][y0(crop) + y][x0(crop) + x][c]
// rate = max(1, F / T): the reference tiles the frame indices until T*rate*views of them exist and takes
// every rate-th one; crop offsets: centre = ceil((dim - S) / 2); with three crops the longer side
// gets 0 / centre / dim - S.  One thread copies 4 output bytes (a row of S*3 bytes is a contiguous
// run of the input row; its start is not 4-byte aligned in general, hence byte loads).
struct Views {
  int F, H, W, T, views, crops, S;
  int y0[3], x0[3];
  int rate;
};

__global__ void __launch_bounds__(256)
eval_views_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const Views v) {
  const int64_t row_bytes = static_cast<int64_t>(v.S) * 3;
  const int64_t rows = static_cast<int64_t>(v.crops) * v.views * v.T * v.S;
  const int64_t total = rows * row_bytes;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
  for (int64_t e = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e < total; e += stride) {
    uint32_t w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t o = e + b;
      if (o >= total) break;
      const int64_t row = o / row_bytes;
      const int xb = static_cast<int>(o - row * row_bytes);
      const int y = static_cast<int>(row % v.S);
      const int64_t r2 = row / v.S;
      const int t = static_cast<int>(r2 % v.T);
      const int64_t r3 = r2 / v.T;
      const int view = static_cast<int>(r3 % v.views), crop = static_cast<int>(r3 / v.views);
      const int f = static_cast<int>((static_cast<int64_t>(view) * v.T + t) * v.rate % v.F);
      const int64_t src = ((static_cast<int64_t>(f) * v.H + v.y0[crop] + y) * v.W + v.x0[crop]) * 3 + xb;
      w |= static_cast<uint32_t>(__ldg(in + src)) << (8 * b);
    }
    if (e + 4 <= total) *reinterpret_cast<uint32_t*>(out + e) = w;
    else for (int b = 0; e + b < total; ++b) out[e + b] = static_cast<uint8_t>(w >> (8 * b));
  }
}

}  // namespace io
}  // namespace x3d

using namespace x3d;

extern "C" int x3d_eval_views_u8(const uint8_t* video, uint8_t* out, int F, int H, int W, int T, int views,
                                 int crops, int S, void* stream) {
  X3D_REQUIRE(video && out, X3D_ERR_INVALID_ARG, "x3d_eval_views_u8: null pointer");
  X3D_REQUIRE(F > 0 && T > 0 && views > 0 && S > 0 && H >= S && W >= S, X3D_ERR_INVALID_ARG,
              "x3d_eval_views_u8: bad extents F=%d T=%d views=%d S=%d H=%d W=%d", F, T, views, S, H, W);
  X3D_REQUIRE(crops == 1 || crops == 3, X3D_ERR_INVALID_ARG, "x3d_eval_views_u8: crops=%d (1 or 3)", crops);
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(out) & 3) == 0, X3D_ERR_INVALID_ARG, "x3d_eval_views_u8: out must be 4-byte aligned");
  io::Views v;
  v.F = F; v.H = H; v.W = W; v.T = T; v.views = views; v.crops = crops; v.S = S;
  v.rate = F / T > 1 ? F / T : 1;
  const int yc = (H - S + 1) / 2, xc = (W - S + 1) / 2;                // ceil((dim - S) / 2)
  for (int i = 0; i < 3; ++i) {
    const int idx = crops > 1 ? i % 3 : 1;                             // left/centre/right vs centre
    v.y0[i] = yc; v.x0[i] = xc;
    if (H > W) { if (idx == 0) v.y0[i] = 0; else if (idx == 2) v.y0[i] = H - S; }
    else       { if (idx == 0) v.x0[i] = 0; else if (idx == 2) v.x0[i] = W - S; }
  }
  const int64_t total = (int64_t)crops * views * T * S * S * 3;
  int64_t blocks = (total / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  io::eval_views_u8_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(video, out, v);
  return check_launch("x3d_eval_views_u8");
}

extern "C" int x3d_normalize_u8(const uint8_t* in, void* out, int64_t pixels, const float* mean,
                                const float* std, float norm_value, int dtype, void* stream) {
  X3D_REQUIRE(in && out && mean && std, X3D_ERR_INVALID_ARG, "x3d_normalize_u8: null pointer");
  X3D_REQUIRE(pixels > 0, X3D_ERR_INVALID_ARG, "x3d_normalize_u8: pixels=%lld", (long long)pixels);
  X3D_REQUIRE(dtype == X3D_F32 || dtype == X3D_BF16, X3D_ERR_INVALID_ARG, "x3d_normalize_u8: dtype %d", dtype);
  X3D_REQUIRE(norm_value != 0.f && std[0] != 0.f && std[1] != 0.f && std[2] != 0.f, X3D_ERR_INVALID_ARG,
              "x3d_normalize_u8: zero divisor");
  X3D_REQUIRE((reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              X3D_ERR_INVALID_ARG, "x3d_normalize_u8: in must be 4-byte, out 16-byte aligned");
  io::Norm nrm;
  for (int i = 0; i < 3; ++i) { nrm.mean[i] = mean[i]; nrm.std[i] = std[i]; }
  nrm.nv = norm_value;
  const int64_t groups = (pixels + 3) / 4;
  int64_t blocks = (groups + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == X3D_F32)
    io::normalize_u8_kernel<float><<<(int)blocks, 256, 0, st>>>(in, static_cast<float*>(out), pixels, nrm);
  else
    io::normalize_u8_kernel<bf16><<<(int)blocks, 256, 0, st>>>(in, static_cast<bf16*>(out), pixels, nrm);
  return check_launch("x3d_normalize_u8");
}

extern "C" int x3d_eval_metrics(const float* probs, const int32_t* labels, double* acc, int V, int ncls,
                                int k, void* stream) {
  X3D_REQUIRE(probs && labels && acc, X3D_ERR_INVALID_ARG, "x3d_eval_metrics: null pointer");
  X3D_REQUIRE(V > 0 && ncls > 0 && k > 0, X3D_ERR_INVALID_ARG, "x3d_eval_metrics: V=%d ncls=%d k=%d", V, ncls, k);
  const int blocks = (V + 3) / 4;
  io::eval_metrics_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(probs, labels, acc, V, ncls, k);
  return check_launch("x3d_eval_metrics");
}
