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
// ties to take and Σ over the selected set.
//
// ONE launch (round 2; round 1 queued six kernels that returned at once in the
// common case): a cooperative grid of at most two 1024-thread CTAs per SM.  Every
// CTA derives the branch of every segment from the forward's counters (a pure
// function of them), CTA s publishes the result of segment s, and when no
// segment needs the fall-back the grid exits — no grid barrier is executed.
// Otherwise the four passes run as loops over the same virtual blocks the
// separate kernels used, separated by grid.sync().  No host synchronisation.
#include <cooperative_groups.h>

#include "common.cuh"

namespace mdseg {
namespace {

constexpr int kBins = 2048;
constexpr int kPasses = 3;
constexpr int kThreads = 1024;
constexpr int kRep = 4;  // replicas of the CTA's histogram (warp w counts into replica w % kRep): hot bins are hot for every warp

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

// branch of a segment: a pure function of the counters the forward left (ohem_ce_loss.py:25,31)
__device__ __forceinline__ bool needs_topk(const mdseg_ohem_state* st) { return st->n_hard < st->n_valid / 16ull; }

// one CTA per segment
__device__ void decide_segment(mdseg_ohem_state* states, unsigned* ws, float* loss_out, int* err_flag, int seg) {
  mdseg_ohem_state* st = states + seg;
  unsigned* H = ws + (size_t)seg * kPasses * kBins;
  if (needs_topk(st))
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
      if (loss_out) loss_out[seg] = loss;
    } else {
      st->mode = 1u;
      st->sum_sel = 0.0;
      st->n_gt = 0ull;
      if (n_min > st->n_px && err_flag) atomicOr(err_flag, MDSEG_ERR_TOPK_RANGE);  // torch.topk would raise
    }
  }
}

// Histogram pass PASS over a run of consecutive images of one segment (pixels [p0, p0 + px_per_image) of loss_px):
// this CTA's interleaved share (vbx of nbx).  A run, not an image, is the unit: 148 CTAs x 1024 threads x 8 values
// sweep 1.2 M values at a time, and per 2 M-pixel image the ragged last sweep left a third of the CTAs waiting at the
// grid barrier.
template <int PASS>
__device__ void radix_hist_block(const float* __restrict__ loss_px, int64_t p0, int64_t px_per_image, int seg,
                                 const mdseg_ohem_state* states, int n_segs, unsigned* ws, int vbx, int nbx,
                                 unsigned* sh /*[kRep * kBins] shared*/) {
  if (seg < 0 || seg >= n_segs) return;
  const mdseg_ohem_state* st = states + seg;
  if (!needs_topk(st)) return;
  unsigned* H = ws + (size_t)seg * kPasses * kBins;

  unsigned* mine = sh + ((threadIdx.x >> 5) & (kRep - 1)) * kBins;
  __shared__ unsigned long long res[3];
  __syncthreads();  // the previous virtual block of this CTA has flushed sh
  for (int i = threadIdx.x; i < kRep * kBins; i += blockDim.x) sh[i] = 0u;
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

  // 16-byte loads, four per thread in flight (a scalar loop keeps 8 KB per SM in flight: latency-bound at 0.8 TB/s);
  // equal digits inside a warp are counted once (match.any): confident nets put most losses into a handful of bins,
  // and same-address shared-memory atomics serialise.
  const float* src = loss_px + p0;
  const bool vec = ((uintptr_t)src & 15) == 0;
  const int64_t n4 = vec ? px_per_image / 4 : 0;
  const int64_t step = (int64_t)nbx * blockDim.x;
  auto count = [&](float v, bool on) {
    const uint32_t key = float_key(v);
    const bool take = on && (PASS == 0 || prefix_of(key, PASS) == want);
    const unsigned d = take ? digit_of(key, PASS) : 0xffffffffu;
    if (PASS == 0) {  // every value is counted and most share a few digits: one atomic per distinct digit of the warp
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (take && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&mine[d], (unsigned)__popc(peers));
    } else if (take) {  // only the values inside the bucket found so far: few, and spread over its sub-digits
      atomicAdd(&mine[d], 1u);
    }
  };
  constexpr int kV = 4;  // 16-byte loads in flight per thread
  for (int64_t i0 = (int64_t)vbx * blockDim.x; i0 < n4; i0 += kV * step) {  // warp-uniform trip count
    float4 v[kV];
    bool on[kV];
#pragma unroll
    for (int u = 0; u < kV; ++u) {
      const int64_t i = i0 + threadIdx.x + u * step;
      on[u] = i < n4;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on[u]) v[u] = __ldcs(reinterpret_cast<const float4*>(src) + i);
    }
#pragma unroll
    for (int u = 0; u < kV; ++u) {
      count(v[u].x, on[u]); count(v[u].y, on[u]); count(v[u].z, on[u]); count(v[u].w, on[u]);
    }
  }
  for (int64_t i0 = 4 * n4 + (int64_t)vbx * blockDim.x; i0 < px_per_image; i0 += step) {  // tail / unaligned images
    const int64_t i = i0 + threadIdx.x;
    const bool on = i < px_per_image;
    count(on ? src[i] : 0.f, on);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kBins; i += blockDim.x)
  {
    unsigned tot = 0;
#pragma unroll
    for (int r = 0; r < kRep; ++r) tot += sh[r * kBins + i];
    if (tot) atomicAdd(H + PASS * kBins + i, tot);
  }
}

// Σ loss over entries strictly above the k-th value (same run / share as radix_hist_block).
__device__ void radix_sum_block(float* __restrict__ loss_px, int64_t p0, int64_t px_per_image, int seg,
                                mdseg_ohem_state* states, int n_segs, const unsigned* ws, int vbx, int nbx) {
  if (seg < 0 || seg >= n_segs) return;
  mdseg_ohem_state* st = states + seg;
  if (!needs_topk(st)) return;
  const unsigned* H = ws + (size_t)seg * kPasses * kBins;
  __shared__ unsigned long long res[3];
  __shared__ double s_sum;
  __shared__ unsigned s_cnt;
  __syncthreads();  // the previous virtual block of this CTA is done with res / s_sum / s_cnt
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
  float* src = loss_px + p0;
  const bool vec = ((uintptr_t)src & 15) == 0;
  const int64_t n4 = vec ? px_per_image / 4 : 0;
  const int64_t step = (int64_t)nbx * blockDim.x;
  // hand out the ties first-come (torch.topk leaves the tie order unspecified too); losers are demoted so that the
  // backward's `loss >= kth` test is exact and needs no atomics.  Returns the value to keep at this position.
  auto visit = [&](float v) -> float {
    const uint32_t key = float_key(v);
    if (key > kkey) { sum += (double)v; ++cnt; }
    else if (key == kkey && atomicAdd(&st->tie_taken, 1u) >= quota) return demoted;
    return v;
  };
  for (int64_t i = (int64_t)vbx * blockDim.x + threadIdx.x; i < n4; i += step) {
    float4* p = reinterpret_cast<float4*>(src) + i;
    const float4 a = *p;
    float4 r;
    r.x = visit(a.x); r.y = visit(a.y); r.z = visit(a.z); r.w = visit(a.w);
    if (r.x != a.x || r.y != a.y || r.z != a.z || r.w != a.w) *p = r;
  }
  for (int64_t i = 4 * n4 + (int64_t)vbx * blockDim.x + threadIdx.x; i < px_per_image; i += step) {
    const float v = src[i];
    const float r = visit(v);
    if (r != v) src[i] = r;
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
__device__ void final_segment(mdseg_ohem_state* states, const unsigned* ws, float* loss_out, int seg) {
  mdseg_ohem_state* st = states + seg;
  if (!needs_topk(st)) return;
  const unsigned* H = ws + (size_t)seg * kPasses * kBins;
  __shared__ unsigned long long res[3];
  const unsigned long long ktot = topk_k(st);
  __syncthreads();
  if (ktot == 0ull) {  // nothing to select (cannot happen unless the segment is empty)
    if (threadIdx.x == 0) {
      st->n_sel = 0; st->inv_n_sel = 0.f; st->loss = __int_as_float(0x7fc00000);
      if (loss_out) loss_out[seg] = st->loss;
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
    if (loss_out) loss_out[seg] = st->loss;
  }
}

struct SelectArgs {
  float* loss_px;
  int64_t px_per_image;
  const int32_t* image_seg;
  mdseg_ohem_state* states;
  unsigned* ws;
  float* loss_out;
  int* err_flag;
  int n_images, n_segs;
};

__global__ void __launch_bounds__(kThreads) ohem_select_kernel(const SelectArgs a) {
  namespace cg = cooperative_groups;
  // which branch: every CTA evaluates every segment; nothing has been written to the states yet
  bool any_topk = false;
  for (int s = 0; s < a.n_segs; ++s) any_topk |= needs_topk(a.states + s);
  __syncthreads();  // all threads have read the counters before thread 0 of a deciding CTA writes next to them
  for (int s = blockIdx.x; s < a.n_segs; s += gridDim.x) decide_segment(a.states, a.ws, a.loss_out, a.err_flag, s);
  if (!any_topk || a.n_images == 0 || a.px_per_image == 0) {
    if (any_topk)  // empty loss vector in the fall-back branch: finish with k = 0
      for (int s = blockIdx.x; s < a.n_segs; s += gridDim.x) final_segment(a.states, a.ws, a.loss_out, s);
    return;  // uniform over the grid: no CTA reaches a grid barrier
  }
  __shared__ unsigned sh[kRep * kBins];  // the CTA's histogram replicas, reused by the three passes
  cg::grid_group grid = cg::this_grid();
  grid.sync();  // histograms zeroed, n_min / mode published
  // runs of consecutive images of one segment (all images when image_seg == NULL): every CTA takes its interleaved
  // share of every run
#define MDSEG_FOR_RUNS(BODY)                                                                  \
  for (int i0 = 0; i0 < a.n_images;) {                                                        \
    const int seg = a.image_seg ? a.image_seg[i0] : 0;                                        \
    int i1 = i0 + 1;                                                                          \
    while (i1 < a.n_images && (a.image_seg ? a.image_seg[i1] : 0) == seg) ++i1;               \
    const int64_t p0 = (int64_t)i0 * a.px_per_image, npx = (int64_t)(i1 - i0) * a.px_per_image; \
    BODY;                                                                                     \
    i0 = i1;                                                                                  \
  }
  MDSEG_FOR_RUNS(radix_hist_block<0>(a.loss_px, p0, npx, seg, a.states, a.n_segs, a.ws, blockIdx.x, gridDim.x, sh));
  grid.sync();
  MDSEG_FOR_RUNS(radix_hist_block<1>(a.loss_px, p0, npx, seg, a.states, a.n_segs, a.ws, blockIdx.x, gridDim.x, sh));
  grid.sync();
  MDSEG_FOR_RUNS(radix_hist_block<2>(a.loss_px, p0, npx, seg, a.states, a.n_segs, a.ws, blockIdx.x, gridDim.x, sh));
  grid.sync();
  MDSEG_FOR_RUNS(radix_sum_block(a.loss_px, p0, npx, seg, a.states, a.n_segs, a.ws, blockIdx.x, gridDim.x));
#undef MDSEG_FOR_RUNS
  grid.sync();
  for (int s = blockIdx.x; s < a.n_segs; s += gridDim.x) final_segment(a.states, a.ws, a.loss_out, s);
}

// resident CTAs of ohem_select_kernel per SM (cooperative launches must fit the chip)
int select_ctas_per_sm() {
  static int cached = 0;
  if (cached == 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ohem_select_kernel, kThreads, 0) != cudaSuccess || n < 1) n = 1;
    cached = n > 2 ? 2 : n;
  }
  return cached;
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
  MDSEG_REQUIRE(n_images == 0 || px_per_image == 0 || loss_px, "mdseg_ohem_select: loss_px is null");
  int64_t bx = ceil_div64(px_per_image, (int64_t)kThreads * 8);
  int64_t want = ceil_div64((int64_t)sm_count() * 2, n_images > 0 ? n_images : 1);
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  SelectArgs a;
  a.loss_px = loss_px; a.px_per_image = px_per_image; a.image_seg = image_seg; a.states = states;
  a.ws = (unsigned*)workspace; a.loss_out = loss_out; a.err_flag = err_flag;
  a.n_images = n_images; a.n_segs = n_segments;
  // CTAs: one per 8 K values of the batch, at most the resident capacity of the chip (cooperative launch)
  int64_t grid = bx * (n_images > 0 ? n_images : 1);
  if (grid < n_segments) grid = n_segments;
  const int64_t cap = (int64_t)sm_count() * select_ctas_per_sm();
  if (grid > cap) grid = cap;
  void* params[] = {(void*)&a};
  MDSEG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)ohem_select_kernel, dim3((unsigned)grid), dim3(kThreads), params, 0, s));
  return 0;
}
