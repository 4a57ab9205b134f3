// eval.cu — evaluator tail on the device.
//
// Reference work replaced (evaluate.py:136-181, MscEvalV0_Contrast.__call__):
//   logits = F.interpolate(logits, size=label HxW, mode='bilinear', align_corners=True)   :149-151
//   probs += torch.softmax(logits, dim=1)                                                  :164
//   (flip) logits = torch.flip(net(flip(im)), dims=(3,)) ; interpolate ; probs += softmax  :165-171
//   preds = torch.argmax(probs, dim=1)                                                     :172
//   label = F.interpolate(label.float(), size=(lH,lW), mode='nearest').long()              :156-157
// The upsampled logits are never materialised: each label pixel interpolates its
// 4 low-res corners per class on the fly (two sweeps: online max/sum, then the
// normalised probabilities are accumulated into `probs`).  The dominant traffic
// is the read-modify-write of probs (8*C bytes per pixel and pass).
#include <float.h>

#include "common.cuh"

namespace mdseg {
namespace {

template <typename T>
__global__ void __launch_bounds__(256)
eval_accum_kernel(const T* __restrict__ logits, int C, int h, int w, float* __restrict__ probs, int H, int W,
                  AxisMap ym, AxisMap xm, int flip, int first) {
  const int64_t HW = (int64_t)H * W, hw = (int64_t)h * w;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    const int Y = (int)(p / W), X = (int)(p - (int64_t)Y * W);
    int y0, y1, x0, x1;
    float l0h, l1h, l0w, l1w;
    ym.at(Y, y0, y1, l0h, l1h);
    xm.at(X, x0, x1, l0w, l1w);
    if (flip) { x0 = w - 1 - x0; x1 = w - 1 - x1; }  // torch.flip(logits, dims=(3,)) before interpolate
    const int64_t o00 = (int64_t)y0 * w + x0, o01 = (int64_t)y0 * w + x1;
    const int64_t o10 = (int64_t)y1 * w + x0, o11 = (int64_t)y1 * w + x1;
    float m = -FLT_MAX, s = 0.f;
    for (int c = 0; c < C; ++c) {
      const T* q = logits + (int64_t)c * hw;
      const float z = l0h * (l0w * to_f32<T>(q[o00]) + l1w * to_f32<T>(q[o01])) +
                      l1h * (l0w * to_f32<T>(q[o10]) + l1w * to_f32<T>(q[o11]));
      const float mn = fmaxf(m, z);
      s = s * ex2_approx((m - mn) * kLog2e) + ex2_approx((z - mn) * kLog2e);
      m = mn;
    }
    const float inv = 1.0f / s;
    for (int c = 0; c < C; ++c) {
      const T* q = logits + (int64_t)c * hw;
      const float z = l0h * (l0w * to_f32<T>(q[o00]) + l1w * to_f32<T>(q[o01])) +
                      l1h * (l0w * to_f32<T>(q[o10]) + l1w * to_f32<T>(q[o11]));
      const float pr = ex2_approx((z - m) * kLog2e) * inv;
      float* dst = probs + (int64_t)c * HW + p;
      *dst = first ? pr : (*dst + pr);
    }
  }
}

template <typename L, bool kSmem>
__global__ void __launch_bounds__(256)
argmax_hist_kernel(const float* __restrict__ probs, int C, int64_t n_px, long long* __restrict__ pred,
                   const L* __restrict__ label, const uint8_t* __restrict__ lut, unsigned long long* __restrict__ hist,
                   int ignore, int* err_flag) {
  extern __shared__ unsigned sh_hist[];
  __shared__ uint8_t s_lut[256];
  const int bins = C * C;
  if (lut) s_lut[threadIdx.x] = lut[threadIdx.x];
  if (kSmem && hist)
    for (int i = threadIdx.x; i < bins; i += blockDim.x) sh_hist[i] = 0u;
  __syncthreads();
  int err = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_px; p += (int64_t)gridDim.x * blockDim.x) {
    float best = probs[p];
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float v = probs[(int64_t)c * n_px + p];
      if (v > best) { best = v; arg = c; }
    }
    if (pred) pred[p] = arg;
    if (hist) {
      int l = load_label<L>(label, p);
      if (lut) l = ((unsigned)l < 256u) ? (int)s_lut[l] : -1;
      if (l != ignore) {
        if ((unsigned)l >= (unsigned)C) err |= MDSEG_ERR_LABEL_RANGE;
        else if (kSmem) atomicAdd(&sh_hist[l * C + arg], 1u);
        else atomicAdd(&hist[l * C + arg], 1ull);
      }
    }
  }
  if (err && err_flag) atomicOr(err_flag, err);
  if (kSmem && hist) {
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
      if (sh_hist[i]) atomicAdd(&hist[i], (unsigned long long)sh_hist[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
label_nearest_kernel(const T* __restrict__ in, int Hin, int Win, T* __restrict__ out, int Hout, int Wout, int n,
                     float sy, float sx) {
  const int64_t total = (int64_t)n * Hout * Wout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xo = (int)(i % Wout);
    const int64_t r = i / Wout;
    const int yo = (int)(r % Hout);
    const int64_t b = r / Hout;
    // ATen nearest_neighbor_compute_source_index: min(floor(dst * scale), in - 1)
    int ys = (int)floorf((float)yo * sy);
    int xs = (int)floorf((float)xo * sx);
    ys = ys > Hin - 1 ? Hin - 1 : ys;
    xs = xs > Win - 1 ? Win - 1 : xs;
    out[i] = in[(b * Hin + ys) * Win + xs];
  }
}

int grid_for(int64_t n) {
  int64_t blocks = ceil_div64(n > 0 ? n : 1, 256);
  int64_t cap = (int64_t)sm_count() * 8;
  return (int)(blocks > cap ? cap : blocks);
}

template <typename L>
int launch_argmax(const float* probs, int C, int64_t n_px, int64_t* pred, const void* label, const uint8_t* lut,
                  int64_t* hist, int ignore, int32_t* ef, cudaStream_t s) {
  const size_t smem = (size_t)C * C * 4;
  const bool use_smem = hist && smem <= 96 * 1024;
  if (use_smem) {
    auto k = argmax_hist_kernel<L, true>;
    if (smem > 48 * 1024) MDSEG_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = grid_for(n_px);
    int cap = sm_count() * (smem > 24 * 1024 ? 2 : 8);
    if (blocks > cap) blocks = cap;
    k<<<blocks, 256, smem, s>>>(probs, C, n_px, (long long*)pred, (const L*)label, lut, (unsigned long long*)hist,
                                ignore, ef);
  } else {
    argmax_hist_kernel<L, false><<<grid_for(n_px), 256, 0, s>>>(probs, C, n_px, (long long*)pred, (const L*)label,
                                                                lut, (unsigned long long*)hist, ignore, ef);
  }
  MDSEG_LAUNCH_OK();
  return 0;
}

}  // namespace
}  // namespace mdseg

extern "C" int mdseg_eval_accum(const void* logits, int dtype, int C, int h, int w, float* probs, int H, int W,
                                int flip, int first, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(logits && probs && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "mdseg_eval_accum: bad arguments");
  AxisMap ym{axis_scale(h, H), h}, xm{axis_scale(w, W), w};
  cudaStream_t s = (cudaStream_t)stream;
  const int g = grid_for((int64_t)H * W);
  switch (dtype) {
    case MDSEG_F32: eval_accum_kernel<float><<<g, 256, 0, s>>>((const float*)logits, C, h, w, probs, H, W, ym, xm, flip, first); break;
    case MDSEG_BF16: eval_accum_kernel<__nv_bfloat16><<<g, 256, 0, s>>>((const __nv_bfloat16*)logits, C, h, w, probs, H, W, ym, xm, flip, first); break;
    case MDSEG_F16: eval_accum_kernel<__half><<<g, 256, 0, s>>>((const __half*)logits, C, h, w, probs, H, W, ym, xm, flip, first); break;
    default: set_error("mdseg_eval_accum: unsupported dtype %d", dtype); return 2;
  }
  MDSEG_LAUNCH_OK();
  return 0;
}

extern "C" int mdseg_argmax_hist(const float* probs, int C, int64_t n_px, int64_t* pred, const void* label,
                                 int label_dtype, const uint8_t* lut256, int64_t* hist, int ignore,
                                 int32_t* err_flag, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(probs && C > 0 && n_px >= 0, "mdseg_argmax_hist: bad arguments");
  MDSEG_REQUIRE(!hist || (label && err_flag), "mdseg_argmax_hist: hist needs label and err_flag");
  MDSEG_REQUIRE((int64_t)C * C < (1LL << 31), "mdseg_argmax_hist: C too large");
  if (n_px == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (!hist) return launch_argmax<uint8_t>(probs, C, n_px, pred, nullptr, nullptr, nullptr, ignore, err_flag, s);
  switch (label_dtype) {
    case MDSEG_U8: return launch_argmax<uint8_t>(probs, C, n_px, pred, label, lut256, hist, ignore, err_flag, s);
    case MDSEG_I32: return launch_argmax<int32_t>(probs, C, n_px, pred, label, lut256, hist, ignore, err_flag, s);
    case MDSEG_I64: return launch_argmax<int64_t>(probs, C, n_px, pred, label, lut256, hist, ignore, err_flag, s);
  }
  set_error("mdseg_argmax_hist: unsupported label dtype %d", label_dtype);
  return 2;
}

extern "C" int mdseg_label_nearest(const void* in, int dtype, int Hin, int Win, void* out, int Hout, int Wout,
                                   int n_images, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(in && out && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && n_images >= 0,
                "mdseg_label_nearest: bad arguments");
  if (n_images == 0) return 0;
  const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
  const int g = grid_for((int64_t)n_images * Hout * Wout);
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case MDSEG_U8: label_nearest_kernel<uint8_t><<<g, 256, 0, s>>>((const uint8_t*)in, Hin, Win, (uint8_t*)out, Hout, Wout, n_images, sy, sx); break;
    case MDSEG_I32: label_nearest_kernel<int32_t><<<g, 256, 0, s>>>((const int32_t*)in, Hin, Win, (int32_t*)out, Hout, Wout, n_images, sy, sx); break;
    case MDSEG_I64: label_nearest_kernel<int64_t><<<g, 256, 0, s>>>((const int64_t*)in, Hin, Win, (int64_t*)out, Hout, Wout, n_images, sy, sx); break;
    default: set_error("mdseg_label_nearest: unsupported dtype %d", dtype); return 2;
  }
  MDSEG_LAUNCH_OK();
  return 0;
}
