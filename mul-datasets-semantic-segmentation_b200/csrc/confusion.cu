// confusion.cu — confusion-matrix accumulation fused with the label LUT
// (SURVEY §8 rows a12, a13).
//
// Reference work replaced (evaluate.py:89-93,174-181; rectangular variants
// :631-634,1738-1741,1866-1869; tools/evaluate_city.py:72-75):
//   keep = label != 255
//   hist += np.bincount(label[keep]*C + pred[keep], minlength=C*C).view(C,C)
// i.e. 5 D2H copies + a boolean gather + a single-thread bincount + one H2D per
// image.  Here: one pass over (label, pred) in HBM, a privatised shared-memory
// histogram per CTA (replicated per warp when it is small), per-thread
// run-length aggregation of equal consecutive keys (segmentation maps are
// piecewise constant), and one 64-bit global atomic per non-empty bin per CTA.
// The accumulator is int64 (the reference's float32 hist is exact only below
// 2^24 per cell).  Algorithmic bytes per pixel: sizeof(label) + sizeof(pred).
#include "common.cuh"

namespace mdseg {
namespace {

constexpr int kPx = 8;  // pixels per thread per iteration

// load 8 consecutive labels/preds as ints; out-of-int-range values become -1
template <typename T> struct Load8;
template <> struct Load8<uint8_t> {
  static __device__ __forceinline__ void load(const uint8_t* p, int (&x)[kPx]) {
    int2 r = ldg_stream_v2(p);
    const uint32_t w[2] = {(uint32_t)r.x, (uint32_t)r.y};
#pragma unroll
    for (int i = 0; i < kPx; ++i) x[i] = (w[i >> 2] >> (8 * (i & 3))) & 0xff;
  }
};
template <> struct Load8<int32_t> {
  static __device__ __forceinline__ void load(const int32_t* p, int (&x)[kPx]) {
    int4 a = ldg_stream_v4(p), b = ldg_stream_v4(p + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  }
};
template <> struct Load8<int64_t> {
  static __device__ __forceinline__ void load(const int64_t* p, int (&x)[kPx]) {
    int4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ldg_stream_v4(p + 2 * j);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[2 * j + 0] = (v[j].y == 0) ? v[j].x : -1;
      x[2 * j + 1] = (v[j].w == 0) ? v[j].z : -1;
    }
  }
};

template <bool kSmem>
__device__ __forceinline__ void bump(unsigned* sh, unsigned long long* hist, int key, unsigned cnt) {
  if (kSmem) atomicAdd(sh + key, cnt);
  else atomicAdd(hist + key, (unsigned long long)cnt);
}

// key of one pixel or -1 (ignored / invalid)
__device__ __forceinline__ int make_key(int l, int p, const uint8_t* s_lut, bool has_lut, int Ca, int Cb, int ignore,
                                        int& err) {
  if (has_lut) {
    if ((unsigned)l < 256u) l = s_lut[l];
    else l = -1;
  }
  if (l == ignore) return -1;
  if ((unsigned)l >= (unsigned)Ca) { err |= MDSEG_ERR_LABEL_RANGE; return -1; }
  if ((unsigned)p >= (unsigned)Cb) { err |= MDSEG_ERR_PRED_RANGE; return -1; }
  return l * Cb + p;
}

template <typename L, typename P, bool kSmem, bool kVec>
__global__ void __launch_bounds__(256)
confusion_kernel(const L* __restrict__ label, const P* __restrict__ pred, const uint8_t* __restrict__ lut,
                 unsigned long long* __restrict__ hist, int Ca, int Cb, int ignore, int64_t n, int* err_flag,
                 int replicas) {
  extern __shared__ unsigned sh_hist[];
  __shared__ uint8_t s_lut[256];
  const int bins = Ca * Cb;
  const bool has_lut = lut != nullptr;
  if (has_lut) s_lut[threadIdx.x] = lut[threadIdx.x];
  if (kSmem) {
    for (int i = threadIdx.x; i < bins * replicas; i += blockDim.x) sh_hist[i] = 0u;
  }
  __syncthreads();
  unsigned* my = sh_hist + (kSmem ? ((threadIdx.x >> 5) % replicas) * bins : 0);

  int err = 0;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  if (kVec) {
    const int64_t nvec = n / kPx;
    for (int64_t v = gtid; v < nvec; v += gstride) {
      int l[kPx], p[kPx];
      Load8<L>::load(label + v * kPx, l);
      Load8<P>::load(pred + v * kPx, p);
      int prev = -1;
      unsigned cnt = 0;
#pragma unroll
      for (int i = 0; i < kPx; ++i) {
        int key = make_key(l[i], p[i], s_lut, has_lut, Ca, Cb, ignore, err);
        if (key == prev) {
          ++cnt;
        } else {
          if (prev >= 0) bump<kSmem>(my, hist, prev, cnt);
          prev = key;
          cnt = 1;
        }
      }
      if (prev >= 0) bump<kSmem>(my, hist, prev, cnt);
    }
    const int64_t t = nvec * kPx + gtid;  // ragged tail (< 8 px)
    if (t < n) {
      int key = make_key(load_label<L>(label, t), load_label<P>(pred, t), s_lut, has_lut, Ca, Cb, ignore, err);
      if (key >= 0) bump<kSmem>(my, hist, key, 1u);
    }
  } else {
    for (int64_t i = gtid; i < n; i += gstride) {
      int key = make_key(load_label<L>(label, i), load_label<P>(pred, i), s_lut, has_lut, Ca, Cb, ignore, err);
      if (key >= 0) bump<kSmem>(my, hist, key, 1u);
    }
  }
  if (err) atomicOr(err_flag, err);

  if (kSmem) {
    __syncthreads();
    for (int b = threadIdx.x; b < bins; b += blockDim.x) {
      unsigned long long s = 0;
      for (int r = 0; r < replicas; ++r) s += sh_hist[r * bins + b];
      if (s) atomicAdd(hist + b, s);
    }
  }
}

// ---- all datasets of a batch in one launch --------------------------------------------------------------
// blockIdx.y = image b; d = dataset_ids[b] selects the LUT (luts + 256*d), the class count C[d] and the
// square histogram hist + offset[d].  Shared-memory privatisation is decided per CTA from C[d].
template <typename L, typename P>
__global__ void __launch_bounds__(256)
confusion_images_kernel(const L* __restrict__ label, const P* __restrict__ pred, const uint8_t* __restrict__ luts,
                        const int32_t* __restrict__ dataset_ids, int64_t px_per_image,
                        unsigned long long* __restrict__ hist, const mdseg_hist_table tab, int ignore,
                        int* err_flag, int smem_words) {
  extern __shared__ unsigned sh_hist[];
  __shared__ uint8_t s_lut[256];
  const int b = blockIdx.y;
  const int d = dataset_ids ? dataset_ids[b] : 0;
  if (d < 0 || d >= tab.n_datasets) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err_flag, MDSEG_ERR_DATASET_ID);
    return;
  }
  const int C = tab.C[d];
  const int bins = C * C;
  unsigned long long* h = hist + tab.offset[d];
  const bool has_lut = luts != nullptr;
  if (has_lut) s_lut[threadIdx.x] = luts[(int64_t)d * 256 + threadIdx.x];
  const bool use_smem = bins <= smem_words;
  int replicas = 1;
  if (use_smem) {
    replicas = smem_words / bins;
    if (replicas > 8) replicas = 8;
    for (int i = threadIdx.x; i < bins * replicas; i += blockDim.x) sh_hist[i] = 0u;
  }
  __syncthreads();
  unsigned* my = sh_hist + (use_smem ? ((threadIdx.x >> 5) % replicas) * bins : 0);
  label += (int64_t)b * px_per_image;
  pred += (int64_t)b * px_per_image;

  int err = 0;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nvec = px_per_image / kPx;  // the host guarantees 16-byte aligned image slices
  for (int64_t v = gtid; v < nvec; v += gstride) {
    int l[kPx], p[kPx];
    Load8<L>::load(label + v * kPx, l);
    Load8<P>::load(pred + v * kPx, p);
    int prev = -1;
    unsigned cnt = 0;
#pragma unroll
    for (int i = 0; i < kPx; ++i) {
      int key = make_key(l[i], p[i], s_lut, has_lut, C, C, ignore, err);
      if (key == prev) {
        ++cnt;
      } else {
        if (prev >= 0) { if (use_smem) atomicAdd(my + prev, cnt); else atomicAdd(h + prev, (unsigned long long)cnt); }
        prev = key;
        cnt = 1;
      }
    }
    if (prev >= 0) { if (use_smem) atomicAdd(my + prev, cnt); else atomicAdd(h + prev, (unsigned long long)cnt); }
  }
  if (err) atomicOr(err_flag, err);
  if (use_smem) {
    __syncthreads();
    for (int k = threadIdx.x; k < bins; k += blockDim.x) {
      unsigned long long sum = 0;
      for (int r = 0; r < replicas; ++r) sum += sh_hist[r * bins + k];
      if (sum) atomicAdd(h + k, sum);
    }
  }
}

// iou / mIoU of every dataset: one CTA per dataset.  Row / column sums are coalesced (a warp walks a row,
// lanes walk columns); the nanmean is a fixed-shape shuffle tree in double, so it is run-to-run deterministic.
__global__ void __launch_bounds__(256) miou_images_kernel(const long long* __restrict__ hist, const mdseg_hist_table tab,
                                                          float* __restrict__ iou, int iou_stride,
                                                          float* __restrict__ miou) {
  extern __shared__ long long s_sum[];  // [2][C]: row sums, column sums
  const int d = blockIdx.x;
  const int C = tab.C[d];
  const long long* h = hist + tab.offset[d];
  float* io = iou + (int64_t)d * iou_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  long long* rows = s_sum;
  long long* cols = s_sum + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) cols[c] = 0;
  __syncthreads();
  for (int r = warp; r < C; r += n_warps) {  // warp r-th row: coalesced; column partials kept per lane
    long long acc = 0;
    for (int j = lane; j < C; j += 32) {
      const long long v = h[(int64_t)r * C + j];
      acc += v;
      if (v) atomicAdd(reinterpret_cast<unsigned long long*>(&cols[j]), (unsigned long long)v);  // integer: exact
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rows[r] = acc;
  }
  __syncthreads();
  double part = 0.0;
  int cnt = 0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const long long dg = h[(int64_t)c * C + c];
    // evaluate.py:96: diag / (sum0 + sum1 - diag); 0/0 -> NaN (class absent)
    const float v = (float)((double)dg / (double)(cols[c] + rows[c] - dg));
    io[c] = v;
    if (v == v) { part += (double)v; ++cnt; }
  }
  __shared__ double s_part[8];
  __shared__ int s_cnt[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    part += __shfl_xor_sync(0xffffffffu, part, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) { s_part[warp] = part; s_cnt[warp] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    int n = 0;
    for (int w = 0; w < n_warps; ++w) { sum += s_part[w]; n += s_cnt[w]; }
    miou[d] = n ? (float)(sum / n) : __int_as_float(0x7fc00000);
  }
}

template <typename L, typename P>
int launch_images(const void* label, const void* pred, const uint8_t* luts, const int32_t* ids, int n_images,
                  int64_t ppi, int64_t* hist, const mdseg_hist_table& tab, int ignore, int32_t* err_flag,
                  cudaStream_t st) {
  int cmax = 0;
  for (int i = 0; i < tab.n_datasets; ++i) cmax = tab.C[i] > cmax ? tab.C[i] : cmax;
  // privatised histogram: up to 24 KB of replicas for small C, one replica up to 96 KB, global atomics beyond
  size_t smem = (size_t)cmax * cmax * 4;
  if (smem < 24 * 1024) smem = 24 * 1024;
  if (smem > 96 * 1024) smem = 24 * 1024;
  auto k = confusion_images_kernel<L, P>;
  if (smem > 48 * 1024) MDSEG_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  int64_t blocks = ceil_div64(ceil_div64(ppi, kPx), 256);
  const int64_t cap = ceil_div64((int64_t)sm_count() * per_sm, n_images);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k<<<dim3((unsigned)blocks, (unsigned)n_images), 256, smem, st>>>(
      (const L*)label, (const P*)pred, luts, ids, ppi, reinterpret_cast<unsigned long long*>(hist), tab, ignore,
      err_flag, (int)(smem / 4));
  MDSEG_LAUNCH_OK();
  return 0;
}

template <typename L>
int dispatch_pred_images(const void* label, const void* pred, int pred_dtype, const uint8_t* luts, const int32_t* ids,
                         int n_images, int64_t ppi, int64_t* hist, const mdseg_hist_table& tab, int ignore,
                         int32_t* err_flag, cudaStream_t st) {
  switch (pred_dtype) {
    case MDSEG_U8: return launch_images<L, uint8_t>(label, pred, luts, ids, n_images, ppi, hist, tab, ignore, err_flag, st);
    case MDSEG_I32: return launch_images<L, int32_t>(label, pred, luts, ids, n_images, ppi, hist, tab, ignore, err_flag, st);
    case MDSEG_I64: return launch_images<L, int64_t>(label, pred, luts, ids, n_images, ppi, hist, tab, ignore, err_flag, st);
  }
  set_error("mdseg_confusion_images: unsupported pred_dtype %d", pred_dtype);
  return 2;
}

template <typename L, typename P>
int launch(const void* label, const void* pred, const uint8_t* lut, int64_t* hist, int Ca, int Cb, int ignore,
           int64_t n, int32_t* err_flag, cudaStream_t st) {
  const int sms = sm_count();
  const int64_t bins = (int64_t)Ca * Cb;
  const bool vec = (((uintptr_t)label | (uintptr_t)pred) & 15) == 0;
  const size_t kMaxSmem = 200 * 1024;
  const bool use_smem = (size_t)bins * 4 <= kMaxSmem;
  int replicas = 1;
  size_t smem = 0;
  int ctas_per_sm = 8;
  if (use_smem) {
    replicas = (int)((24 * 1024) / (bins * 4));
    if (replicas > 8) replicas = 8;
    if (replicas < 1) replicas = 1;
    smem = (size_t)bins * 4 * replicas;
    int fit = (int)((220 * 1024) / (smem + 1024));
    if (fit < 1) fit = 1;
    if (ctas_per_sm > fit) ctas_per_sm = fit;
  }
  int64_t work_threads = vec ? ceil_div64(n, kPx) : n;
  int64_t blocks = ceil_div64(work_threads > 0 ? work_threads : 1, 256);
  int64_t cap = (int64_t)sms * ctas_per_sm;
  if (blocks > cap) blocks = cap;
  MDSEG_REQUIRE(ceil_div64(n, blocks) < (int64_t)0xffffffffLL, "mdseg_confusion: n too large for one launch");
  // err_flag may be NULL: point at a scratch word inside hist? No — require it.
  auto* h = reinterpret_cast<unsigned long long*>(hist);
#define MDSEG_CONF_LAUNCH(SM, VE)                                                                             \
  do {                                                                                                        \
    auto k = confusion_kernel<L, P, SM, VE>;                                                                  \
    if (smem > 48 * 1024) MDSEG_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k<<<(unsigned)blocks, 256, smem, st>>>((const L*)label, (const P*)pred, lut, h, Ca, Cb, ignore, n, err_flag, replicas); \
  } while (0)
  if (use_smem) { if (vec) MDSEG_CONF_LAUNCH(true, true); else MDSEG_CONF_LAUNCH(true, false); }
  else          { if (vec) MDSEG_CONF_LAUNCH(false, true); else MDSEG_CONF_LAUNCH(false, false); }
#undef MDSEG_CONF_LAUNCH
  MDSEG_LAUNCH_OK();
  return 0;
}

template <typename L>
int dispatch_pred(const void* label, const void* pred, int pred_dtype, const uint8_t* lut, int64_t* hist, int Ca,
                  int Cb, int ignore, int64_t n, int32_t* err_flag, cudaStream_t st) {
  switch (pred_dtype) {
    case MDSEG_U8: return launch<L, uint8_t>(label, pred, lut, hist, Ca, Cb, ignore, n, err_flag, st);
    case MDSEG_I32: return launch<L, int32_t>(label, pred, lut, hist, Ca, Cb, ignore, n, err_flag, st);
    case MDSEG_I64: return launch<L, int64_t>(label, pred, lut, hist, Ca, Cb, ignore, n, err_flag, st);
  }
  set_error("mdseg_confusion: unsupported pred_dtype %d", pred_dtype);
  return 2;
}

// iou / mIoU on the device: one CTA; the final nanmean is a sequential double
// sum by one thread so the result is run-to-run deterministic.
__global__ void miou_kernel(const long long* __restrict__ hist, int C, float* __restrict__ iou,
                            float* __restrict__ miou) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    long long row = 0, col = 0;
    for (int j = 0; j < C; ++j) { row += hist[(int64_t)c * C + j]; col += hist[(int64_t)j * C + c]; }
    long long d = hist[(int64_t)c * C + c];
    // evaluate.py:96: diag / (sum0 + sum1 - diag); 0/0 -> NaN (class absent)
    iou[c] = (float)((double)d / (double)(col + row - d));
  }
  __syncthreads();
  if (threadIdx.x == 0 && miou) {
    double s = 0.0;
    int cnt = 0;
    for (int c = 0; c < C; ++c) {
      float v = iou[c];
      if (v == v) { s += (double)v; ++cnt; }
    }
    *miou = cnt ? (float)(s / cnt) : __int_as_float(0x7fc00000);
  }
}

}  // namespace
}  // namespace mdseg

extern "C" int mdseg_confusion(const void* label, int label_dtype, const void* pred, int pred_dtype,
                               const uint8_t* lut256, int64_t* hist, int Ca, int Cb, int ignore, int64_t n,
                               int32_t* err_flag, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(Ca > 0 && Cb > 0 && (int64_t)Ca * Cb < (1LL << 31), "mdseg_confusion: bad Ca/Cb %d %d", Ca, Cb);
  MDSEG_REQUIRE(n >= 0, "mdseg_confusion: n < 0");
  if (n == 0) return 0;
  MDSEG_REQUIRE(label && pred && hist && err_flag, "mdseg_confusion: null pointer (err_flag is required)");
  cudaStream_t st = (cudaStream_t)stream;
  switch (label_dtype) {
    case MDSEG_U8: return dispatch_pred<uint8_t>(label, pred, pred_dtype, lut256, hist, Ca, Cb, ignore, n, err_flag, st);
    case MDSEG_I32: return dispatch_pred<int32_t>(label, pred, pred_dtype, lut256, hist, Ca, Cb, ignore, n, err_flag, st);
    case MDSEG_I64: return dispatch_pred<int64_t>(label, pred, pred_dtype, lut256, hist, Ca, Cb, ignore, n, err_flag, st);
  }
  set_error("mdseg_confusion: unsupported label_dtype %d", label_dtype);
  return 2;
}

extern "C" int mdseg_miou(const int64_t* hist, int C, float* iou, float* miou, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(hist && iou && C > 0, "mdseg_miou: bad arguments (iou is required)");
  miou_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((const long long*)hist, C, iou, miou);
  MDSEG_LAUNCH_OK();
  return 0;
}

extern "C" int mdseg_confusion_images(const void* label, int label_dtype, const void* pred, int pred_dtype,
                                      const uint8_t* luts, const int32_t* dataset_ids, int n_images,
                                      int64_t px_per_image, int64_t* hist, const mdseg_hist_table* tab, int ignore,
                                      int32_t* err_flag, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(tab && tab->n_datasets > 0 && tab->n_datasets <= MDSEG_MAX_DATASETS, "mdseg_confusion_images: bad table");
  for (int i = 0; i < tab->n_datasets; ++i)
    MDSEG_REQUIRE(tab->C[i] > 0 && tab->C[i] < 32768 && tab->offset[i] >= 0, "mdseg_confusion_images: bad table entry %d", i);
  MDSEG_REQUIRE(n_images >= 0 && n_images <= 65535 && px_per_image >= 0, "mdseg_confusion_images: bad shape");
  if (n_images == 0 || px_per_image == 0) return 0;
  MDSEG_REQUIRE(label && pred && hist && err_flag, "mdseg_confusion_images: null pointer (err_flag is required)");
  MDSEG_REQUIRE(px_per_image % 16 == 0 && (((uintptr_t)label | (uintptr_t)pred) & 15) == 0,
                "mdseg_confusion_images: images must be 16-element multiples and 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  switch (label_dtype) {
    case MDSEG_U8: return dispatch_pred_images<uint8_t>(label, pred, pred_dtype, luts, dataset_ids, n_images, px_per_image, hist, *tab, ignore, err_flag, st);
    case MDSEG_I32: return dispatch_pred_images<int32_t>(label, pred, pred_dtype, luts, dataset_ids, n_images, px_per_image, hist, *tab, ignore, err_flag, st);
    case MDSEG_I64: return dispatch_pred_images<int64_t>(label, pred, pred_dtype, luts, dataset_ids, n_images, px_per_image, hist, *tab, ignore, err_flag, st);
  }
  set_error("mdseg_confusion_images: unsupported label_dtype %d", label_dtype);
  return 2;
}

extern "C" int mdseg_miou_images(const int64_t* hist, const mdseg_hist_table* tab, float* iou, int iou_stride,
                                 float* miou, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(hist && tab && iou && miou && tab->n_datasets > 0 && tab->n_datasets <= MDSEG_MAX_DATASETS,
                "mdseg_miou_images: bad arguments");
  for (int i = 0; i < tab->n_datasets; ++i)
    MDSEG_REQUIRE(tab->C[i] > 0 && tab->C[i] <= iou_stride, "mdseg_miou_images: iou_stride %d < C[%d]", iou_stride, i);
  int cmax = 0;
  for (int i = 0; i < tab->n_datasets; ++i) cmax = tab->C[i] > cmax ? tab->C[i] : cmax;
  MDSEG_REQUIRE((size_t)cmax * 16 <= 48 * 1024, "mdseg_miou_images: more than 3072 classes");
  miou_images_kernel<<<tab->n_datasets, 256, (size_t)cmax * 16, (cudaStream_t)stream>>>((const long long*)hist, *tab,
                                                                                       iou, iou_stride, miou);
  MDSEG_LAUNCH_OK();
  return 0;
}
