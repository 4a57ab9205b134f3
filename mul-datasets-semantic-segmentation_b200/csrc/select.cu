// select.cu — OHEM hard-pixel selection without a sort.
//
// Reference work replaced (lib/loss/ohem_ce_loss.py:25-34 and :52,70-90):
//   n_min = labels[labels != 255].numel() // 16          (host sync)
//   loss_hard = loss[loss > thresh]                      (nonzero + index, host sync)
//   if loss_hard.numel() < n_min: loss_hard, _ = loss.topk(n_min)   (full sort/select)
//   return torch.mean(loss_hard)
// Here everything stays on the device: the forward kernels already counted
// n_valid / n_hard / Σ hard; `decide` picks the branch; the fall-back branch is a
// 3-digit (11+11+10 bit) MSD radix select over the order-preserving integer
// image of the fp32 losses — three histogram passes and one summation pass over
// the 4 B/px loss array — which yields the k-th largest value, the number of
// ties to take and Σ over the selected set.  All pass kernels return
// immediately when the threshold branch was taken, so the common case costs six
// empty launches and no host synchronisation.
#include "common.cuh"

namespace mdseg {
namespace {

constexpr int kBins = 2048;
constexpr int kPasses = 3;
constexpr int kThreads = 1024;

__device__ __forceinline__ unsigned long long topk_k(const mdseg_ohem_state* st) {
  unsigned long long k = st->n_min;
  return k > st->n_px ? st->n_px : k;
}

// Block-wide (1024 threads, 2048 bins): the bin b with
//   #{entries in bins > b} < k <= #{entries in bins >= b}.
// Returns through shared memory: res[0] = b, res[1] = k - #{> b}, res[2] = hist[b].
__device__ void find_bucket(const unsigned* __restrict__ hist, unsigned long long k, unsigned long long* res) {
  __shared__ unsigned long long warp_tot[32];
  const int t = threadIdx.x;
  const int lane = t & 31, wid = t >> 5;
  const unsigned h0 = hist[2 * t], h1 = hist[2 * t + 1];
  unsigned long long local = (unsigned long long)h0 + h1;
  // inclusive suffix scan inside the warp (towards higher lanes)
  unsigned long long incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long v = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += v;
  }
  if (lane == 0) warp_tot[wid] = incl;
  __syncthreads();
  unsigned long long above = 0;  // sum over warps with a higher index
  for (int w = wid + 1; w < 32; ++w) above += warp_tot[w];
  const unsigned long long excl = above + incl - local;  // Σ over threads > t
  // bin 2t+1
  unsigned long long gt = excl, ge = excl + h1;
  if (gt < k && k <= ge) { res[0] = 2 * t + 1; res[1] = k - gt; res[2] = h1; }
  gt = ge; ge = gt + h0;  // bin 2t
  if (gt < k && k <= ge) { res[0] = 2 * t; res[1] = k - gt; res[2] = h0; }
  __syncthreads();
}

__device__ __forceinline__ unsigned digit_of(uint32_t key, int pass) {
  return pass == 0 ? (key >> 21) : pass == 1 ? ((key >> 10) & 2047u) : (key & 1023u);
}
__device__ __forceinline__ uint32_t prefix_of(uint32_t key, int pass) {  // bits above this pass' digit
  return pass == 0 ? 0u : pass == 1 ? (key >> 21) : (key >> 10);
}

__global__ void begin_kernel(mdseg_ohem_state* st, int n, float thresh) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mdseg_ohem_state z;
    memset(&z, 0, sizeof(z));
    z.thresh = thresh;
    st[i] = z;
  }
}

// one CTA per segment
__global__ void __launch_bounds__(kThreads)
decide_kernel(mdseg_ohem_state* states, unsigned* ws, float* loss_out, int* err_flag) {
  mdseg_ohem_state* st = states + blockIdx.x;
  unsigned* H = ws + (size_t)blockIdx.x * kPasses * kBins;
  for (int i = threadIdx.x; i < kPasses * kBins; i += blockDim.x) H[i] = 0u;
  if (threadIdx.x == 0) {
    const unsigned long long n_min = st->n_valid / 16ull;  // ohem_ce_loss.py:25
    st->n_min = n_min;
    st->tie_taken = 0u;
    if (st->n_hard >= n_min) {  // ohem_ce_loss.py:31: topk only if numel() < n_min
      st->mode = 0u;
      st->n_sel = st->n_hard;
      st->sum_sel = st->sum_hard;
      const float loss = st->n_hard ? (float)(st->sum_hard / (double)st->n_hard) : __int_as_float(0x7fc00000);
      st->loss = loss;
      st->inv_n_sel = st->n_hard ? (float)(1.0 / (double)st->n_hard) : 0.f;
      if (loss_out) loss_out[blockIdx.x] = loss;
    } else {
      st->mode = 1u;
      st->sum_sel = 0.0;
      st->n_gt = 0ull;
      if (n_min > st->n_px && err_flag) atomicOr(err_flag, MDSEG_ERR_TOPK_RANGE);  // torch.topk would raise
    }
  }
}

template <int PASS>
__global__ void __launch_bounds__(kThreads)
radix_hist_kernel(const float* __restrict__ loss_px, int64_t px_per_image, const int32_t* __restrict__ image_seg,
                  const mdseg_ohem_state* __restrict__ states, int n_segs, unsigned* __restrict__ ws) {
  const int img = blockIdx.y;
  const int seg = image_seg ? image_seg[img] : 0;
  if (seg < 0 || seg >= n_segs) return;
  const mdseg_ohem_state* st = states + seg;
  if (st->mode != 1u) return;
  unsigned* H = ws + (size_t)seg * kPasses * kBins;

  __shared__ unsigned sh[kBins];
  __shared__ unsigned long long res[3];
  for (int i = threadIdx.x; i < kBins; i += blockDim.x) sh[i] = 0u;
  uint32_t want = 0;
  if (PASS >= 1) {
    find_bucket(H, topk_k(st), res);
    want = (uint32_t)res[0];
    unsigned long long k1 = res[1];
    if (PASS == 2) {
      __syncthreads();
      find_bucket(H + kBins, k1, res);
      want = (want << 11) | (uint32_t)res[0];
    }
  }
  __syncthreads();

  const float* src = loss_px + (int64_t)img * px_per_image;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px_per_image;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = float_key(src[i]);
    if (PASS == 0 || prefix_of(key, PASS) == want) atomicAdd(&sh[digit_of(key, PASS)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
    if (sh[i]) atomicAdd(H + PASS * kBins + i, sh[i]);
}

// Σ loss over entries strictly above the k-th value.
__global__ void __launch_bounds__(kThreads)
radix_sum_kernel(float* __restrict__ loss_px, int64_t px_per_image, const int32_t* __restrict__ image_seg,
                 mdseg_ohem_state* __restrict__ states, int n_segs, const unsigned* __restrict__ ws) {
  const int img = blockIdx.y;
  const int seg = image_seg ? image_seg[img] : 0;
  if (seg < 0 || seg >= n_segs) return;
  mdseg_ohem_state* st = states + seg;
  if (st->mode != 1u) return;
  const unsigned* H = ws + (size_t)seg * kPasses * kBins;
  __shared__ unsigned long long res[3];
  __shared__ double s_sum;
  __shared__ unsigned s_cnt;
  find_bucket(H, topk_k(st), res);
  uint32_t kkey = (uint32_t)res[0];
  unsigned long long k = res[1];
  __syncthreads();
  find_bucket(H + kBins, k, res);
  kkey = (kkey << 11) | (uint32_t)res[0];
  k = res[1];
  __syncthreads();
  find_bucket(H + 2 * kBins, k, res);
  kkey = (kkey << 10) | (uint32_t)res[0];
  const unsigned quota = (unsigned)res[1];
  if (threadIdx.x == 0) { s_sum = 0.0; s_cnt = 0u; }
  __syncthreads();
  // value stored over a tie that did not make the quota: just below kth
  const float kth = key_float(kkey);
  const float demoted = kth > 0.f ? __uint_as_float(__float_as_uint(kth) - 1u) : -1.17549435e-38f;

  double sum = 0.0;
  unsigned cnt = 0;
  float* src = loss_px + (int64_t)img * px_per_image;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < px_per_image;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    const uint32_t key = float_key(v);
    if (key > kkey) { sum += (double)v; ++cnt; }
    else if (key == kkey) {
      // hand out the ties first-come (torch.topk leaves the tie order
      // unspecified too); losers are demoted so that the backward's
      // `loss >= kth` test is exact and needs no atomics.
      if (atomicAdd(&st->tie_taken, 1u) >= quota) src[i] = demoted;
    }
  }
  sum = warp_sum(sum);
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0 && cnt) { atomicAdd(&s_sum, sum); atomicAdd(&s_cnt, cnt); }
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) {
    atomicAdd(&st->sum_sel, s_sum);
    atomicAdd(&st->n_gt, (unsigned long long)s_cnt);
  }
}

// one CTA per segment: finish the top-k branch
__global__ void __launch_bounds__(kThreads)
final_kernel(mdseg_ohem_state* states, const unsigned* __restrict__ ws, float* loss_out) {
  mdseg_ohem_state* st = states + blockIdx.x;
  if (st->mode != 1u) return;
  const unsigned* H = ws + (size_t)blockIdx.x * kPasses * kBins;
  __shared__ unsigned long long res[3];
  const unsigned long long ktot = topk_k(st);
  if (ktot == 0ull) {  // nothing to select (cannot happen unless the segment is empty)
    if (threadIdx.x == 0) {
      st->n_sel = 0; st->inv_n_sel = 0.f; st->loss = __int_as_float(0x7fc00000);
      if (loss_out) loss_out[blockIdx.x] = st->loss;
    }
    return;
  }
  find_bucket(H, ktot, res);
  uint32_t kkey = (uint32_t)res[0];
  unsigned long long k = res[1];
  __syncthreads();
  find_bucket(H + kBins, k, res);
  kkey = (kkey << 11) | (uint32_t)res[0];
  k = res[1];
  __syncthreads();
  find_bucket(H + 2 * kBins, k, res);
  kkey = (kkey << 10) | (uint32_t)res[0];
  if (threadIdx.x == 0) {
    const float kth = key_float(kkey);
    const unsigned quota = (unsigned)res[1];
    st->kth = kth;
    st->tie_quota = quota;
    st->n_ties = (unsigned)res[2];
    st->n_sel = ktot;
    const double total = st->sum_sel + (double)quota * (double)kth;
    st->sum_sel = total;
    st->loss = (float)(total / (double)ktot);
    st->inv_n_sel = (float)(1.0 / (double)ktot);
    if (loss_out) loss_out[blockIdx.x] = st->loss;
  }
}

}  // namespace
}  // namespace mdseg

extern "C" size_t mdseg_select_workspace_bytes(int n_segments) {
  return (size_t)(n_segments > 0 ? n_segments : 1) * mdseg::kPasses * mdseg::kBins * sizeof(unsigned);
}

extern "C" int mdseg_ohem_begin(mdseg_ohem_state* states, int n_segments, float thresh, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(states && n_segments > 0, "mdseg_ohem_begin: bad arguments");
  begin_kernel<<<(n_segments + 63) / 64, 64, 0, (cudaStream_t)stream>>>(states, n_segments, thresh);
  MDSEG_LAUNCH_OK();
  return 0;
}

extern "C" int mdseg_ohem_select(float* loss_px, int n_images, int64_t px_per_image, const int32_t* image_seg,
                                 mdseg_ohem_state* states, int n_segments, void* workspace, float* loss_out,
                                 int32_t* err_flag, void* stream) {
  using namespace mdseg;
  MDSEG_REQUIRE(states && workspace && n_segments > 0, "mdseg_ohem_select: null pointer");
  MDSEG_REQUIRE(n_images >= 0 && n_images <= 65535 && px_per_image >= 0, "mdseg_ohem_select: bad image count");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned* ws = (unsigned*)workspace;
  decide_kernel<<<n_segments, kThreads, 0, s>>>(states, ws, loss_out, err_flag);
  MDSEG_LAUNCH_OK();
  if (n_images == 0 || px_per_image == 0) return 0;
  MDSEG_REQUIRE(loss_px, "mdseg_ohem_select: loss_px is null");
  // enough CTAs to fill the chip, at most one per 8 K pixels of an image
  int64_t bx = ceil_div64(px_per_image, (int64_t)kThreads * 8);
  int64_t want = ceil_div64((int64_t)sm_count() * 2, n_images);
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)n_images);
  radix_hist_kernel<0><<<grid, kThreads, 0, s>>>(loss_px, px_per_image, image_seg, states, n_segments, ws);
  radix_hist_kernel<1><<<grid, kThreads, 0, s>>>(loss_px, px_per_image, image_seg, states, n_segments, ws);
  radix_hist_kernel<2><<<grid, kThreads, 0, s>>>(loss_px, px_per_image, image_seg, states, n_segments, ws);
  radix_sum_kernel<<<grid, kThreads, 0, s>>>(loss_px, px_per_image, image_seg, states, n_segments, ws);
  final_kernel<<<n_segments, kThreads, 0, s>>>(states, ws, loss_out);
  MDSEG_LAUNCH_OK();
  return 0;
}
