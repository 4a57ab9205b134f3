// label_pipeline.cu — the label branch of the training data pipeline in one pass.
//
// Reference work replaced, per sample, on a DataLoader worker (CPU, numpy / cv2):
//   label = self.lb_map[label]                                      lib/base_dataset.py:81-82
//   lb = cv2.resize(lb, (im_w, im_h), interpolation=INTER_NEAREST)  lib/transform_cv2.py:43
//   lb = np.pad(lb, ((ph, ph), (pw, pw)), constant_values=255)      lib/transform_cv2.py:52-53
//   lb = lb[sh:sh+crop_h, sw:sw+crop_w]                             lib/transform_cv2.py:57-61
//   lb = lb[:, ::-1]                                                lib/transform_cv2.py:71-77
//   torch.from_numpy(lb.astype(np.int64))                           lib/transform_cv2.py:300
// Every step is an index map, so the whole chain is one gather:
//   out[b, y, x] = lut_b[ src_b[ sy(Y) ][ sx(X) ] ]   with X = (flip ? crop_w-1-x : x) + crop_x - pad_left,
//                                                          Y = y + crop_y - pad_top,
//   255 (the pad value, NOT passed through the LUT) when (Y, X) falls outside the resized image, and
//   s(v) = min(floor(v * (1 / (dst / src))), src - 1) in double: OpenCV's resizeNN for INTER_NEAREST.
// The host draws the random numbers (scale, crop origin, flip) exactly as the reference does and passes the resulting
// integers per image; the raw label images stay uint8 in HBM and never exist at the intermediate sizes.
//
// HBM-bound byte kernel: a thread produces 16 consecutive output pixels of a row (one 16-byte store for uint8 output,
// eight for int64), reads go through L1 (neighbouring outputs read neighbouring or identical source bytes).
// Algorithmic bytes per output pixel: sizeof(out) + (source bytes actually touched) <= sizeof(out) + 1.
#include "common.cuh"

namespace mdseg {
namespace {

constexpr int kPx = 16;  // output pixels per thread

template <typename Out>
__device__ __forceinline__ void store16(Out* p, const int (&v)[kPx]);
template <>
__device__ __forceinline__ void store16<uint8_t>(uint8_t* p, const int (&v)[kPx]) {
  uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < kPx; ++i) w[i >> 2] |= ((uint32_t)v[i] & 0xffu) << (8 * (i & 3));
  *reinterpret_cast<int4*>(p) = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
}
template <>
__device__ __forceinline__ void store16<int64_t>(int64_t* p, const int (&v)[kPx]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) reinterpret_cast<int4*>(p)[j] = make_int4(v[2 * j], 0, v[2 * j + 1], 0);
}

// blockIdx.y = image.  out_w % 16 == 0 is required by the launcher (vector stores); rows are walked by a grid-stride
// loop over (row, 16-pixel group) pairs.
template <typename Out>
__global__ void __launch_bounds__(256) label_pipeline_kernel(const mdseg_label_view* __restrict__ views,
                                                            const uint8_t* __restrict__ luts, int n_luts,
                                                            Out* __restrict__ out, int out_h, int out_w, int pad_value) {
  __shared__ uint8_t s_lut[256];
  const mdseg_label_view v = views[blockIdx.y];
  s_lut[threadIdx.x] = (luts && v.lut >= 0 && v.lut < n_luts) ? luts[(int64_t)v.lut * 256 + threadIdx.x]
                                                              : (uint8_t)threadIdx.x;
  __syncthreads();
  // OpenCV: inv_scale = (double)dsize / ssize; scale = 1. / inv_scale  (resize.cpp, resizeNN)
  const double ifx = 1.0 / ((double)v.im_w / (double)v.src_w);
  const double ify = 1.0 / ((double)v.im_h / (double)v.src_h);
  const int groups = out_w / kPx;
  const int64_t total = (int64_t)out_h * groups;
  Out* outb = out + (int64_t)blockIdx.y * out_h * out_w;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(t / groups), x0 = (int)(t - (int64_t)y * groups) * kPx;
    const int Y = y + v.crop_y - v.pad_top;
    int val[kPx];
    if (Y < 0 || Y >= v.im_h) {
#pragma unroll
      for (int i = 0; i < kPx; ++i) val[i] = pad_value;
    } else {
      int sy = (int)floor((double)Y * ify);
      sy = sy > v.src_h - 1 ? v.src_h - 1 : sy;
      const uint8_t* row = v.src + (int64_t)sy * v.src_row_stride;
#pragma unroll
      for (int i = 0; i < kPx; ++i) {
        const int xo = x0 + i;
        const int X = (v.flip ? out_w - 1 - xo : xo) + v.crop_x - v.pad_left;
        int r = pad_value;
        if (X >= 0 && X < v.im_w) {
          int sx = (int)floor((double)X * ifx);
          sx = sx > v.src_w - 1 ? v.src_w - 1 : sx;
          r = s_lut[__ldg(row + sx)];
        }
        val[i] = r;
      }
    }
    store16<Out>(outb + (int64_t)y * out_w + x0, val);
  }
}

}  // namespace
}  // namespace mdseg

extern "C" int mdseg_label_pipeline(const mdseg_label_view* views, int n_images, const uint8_t* luts, int n_luts,
                                    void* out, int out_dtype, int out_h, int out_w, int pad_value, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(views && out && n_images >= 0 && out_h > 0 && out_w > 0, "mdseg_label_pipeline: bad arguments");
  MDSEG_REQUIRE(out_w % kPx == 0, "mdseg_label_pipeline: out_w (%d) must be a multiple of 16", out_w);
  MDSEG_REQUIRE(pad_value >= 0 && pad_value <= 255, "mdseg_label_pipeline: pad_value outside [0, 255]");
  MDSEG_REQUIRE(n_luts >= 0 && (n_luts == 0 || luts), "mdseg_label_pipeline: n_luts without a table");
  if (n_images == 0) return 0;
  const int64_t threads = (int64_t)out_h * (out_w / kPx);
  int64_t blocks = ceil_div64(threads, 256);
  const int64_t cap = ceil_div64((int64_t)sm_count() * 8, n_images);  // eight resident CTAs per SM over all images
  if (blocks > cap) blocks = cap;
  const dim3 grid((unsigned)blocks, (unsigned)n_images);
  cudaStream_t s = (cudaStream_t)stream;
  switch (out_dtype) {
    case MDSEG_U8:
      label_pipeline_kernel<uint8_t><<<grid, 256, 0, s>>>(views, luts, n_luts, (uint8_t*)out, out_h, out_w, pad_value);
      break;
    case MDSEG_I64:
      label_pipeline_kernel<int64_t><<<grid, 256, 0, s>>>(views, luts, n_luts, (int64_t*)out, out_h, out_w, pad_value);
      break;
    default:
      set_error("mdseg_label_pipeline: output dtype %d (uint8 or int64)", out_dtype);
      return 2;
  }
  MDSEG_LAUNCH_OK();
  return 0;
}
