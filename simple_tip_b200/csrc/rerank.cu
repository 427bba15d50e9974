// rerank.cu — exact stage of the nearest-neighbour search (surprise.py:633-651).
//
// For every query row the tensor-core filter (pair_tc.cu) leaves a candidate set that provably
// contains NumPy's argmin.  Here each candidate's distance is recomputed exactly as
//   np.linalg.norm(from[:, None] - to, axis=2)   (surprise.py:638-640)
// does it: difference, square, NumPy's pairwise summation order, correctly rounded sqrt — all in
// the input dtype with no FMA contraction — and the winner is the lexicographic minimum of
// (distance, original train index), i.e. np.argmin's first occurrence (surprise.py:645-647).
// Rows without a usable candidate list (0 or > cap candidates, or no filter at all) are scanned
// exhaustively over their class range, which is also the reference-grade fallback path.
#include <algorithm>
#include "common.cuh"

namespace tip {

constexpr int kRerankThreads = 128;
constexpr int kCandGroup = 8;   // must match the group size of the filter epilogue (pair_tc.cu)

template <typename T>
struct Best {
  T dist;
  int gid;
  int pos;
};

template <typename T>
__device__ __forceinline__ void consider(Best<T>& b, T dist, int gid, int pos) {
  // NaN distances never win (np.min would propagate NaN; documented as unsupported input)
  if (dist < b.dist || (dist == b.dist && gid < b.gid)) {
    b.dist = dist;
    b.gid = gid;
    b.pos = pos;
  }
}

template <typename T>
__global__ void __launch_bounds__(kRerankThreads)
rerank_kernel(const T* __restrict__ q, const T* __restrict__ t, int64_t m, int64_t n, int d,
              const int32_t* __restrict__ cand_idx, const int32_t* __restrict__ cand_cnt, int cap,
              const int32_t* __restrict__ q_class, const int32_t* __restrict__ class_off, int n_classes,
              int mode, const int32_t* __restrict__ t_gid, T* __restrict__ out_dist,
              int32_t* __restrict__ out_pos, unsigned long long* __restrict__ stats, bool xs_ok) {
  __shared__ T s_dist[kRerankThreads];
  __shared__ int s_gid[kRerankThreads];
  __shared__ int s_pos[kRerankThreads];
  extern __shared__ __align__(16) unsigned char rerank_smem[];
  T* xs = reinterpret_cast<T*>(rerank_smem);
  for (int64_t row = blockIdx.x; row < m; row += gridDim.x) {
    const T* xg = q + row * (int64_t)d;
    const T* x = xg;
    if (xs_ok) {                       // query row reused by every candidate: keep it on chip
      for (int i = threadIdx.x; i < d; i += kRerankThreads) xs[i] = xg[i];
      __syncthreads();
      x = xs;
    }
    Best<T> best;
    best.dist = Rn<T>::inf();
    best.gid = 0x7fffffff;
    best.pos = -1;
    const int cnt = cand_cnt ? cand_cnt[row] : 0;
    const int cls = q_class ? q_class[row] : 0;
    const bool listed = cand_cnt != nullptr && cnt >= 1 && cnt <= cap;
    const int sub = threadIdx.x & 7;            // lane inside its 8-lane group
    const int grp = threadIdx.x >> 3;           // 16 groups per block, one train row each
    constexpr int kGroups = kRerankThreads / 8;
    if (listed && cls >= 0 && cls < n_classes) {
      // every candidate names a group of kCandGroup consecutive train rows (clipped to the
      // class range it starts in)
      const int c0 = class_off[cls], c1 = class_off[cls + 1], cn = class_off[n_classes];
      const int total = cnt * kCandGroup;
      for (int base = 0; base < total; base += kGroups) {
        const int k = base + grp;
        bool ok = k < total;
        int j = 0;
        if (ok) {
          const int start = cand_idx[row * (int64_t)cap + k / kCandGroup];
          j = start + (k % kCandGroup);
          const int limit = mode == TIP_RANGE_SAME_CLASS ? c1 : (start < c0 ? c0 : cn);
          ok = start >= 0 && j < limit && j < n;
          if (!ok) j = 0;
        }
        const T s = np_sumsq_g8<T>(x, t + (int64_t)j * d, d, sub);
        if (ok && sub == 0) consider(best, Rn<T>::sqrt(s), t_gid ? t_gid[j] : j, j);
      }
      if (threadIdx.x == 0 && stats) atomicAdd(stats + 1, (unsigned long long)cnt);
    } else if (cls >= 0 && cls < n_classes && n > 0) {
      const int c0 = class_off[cls], c1 = class_off[cls + 1], cn = class_off[n_classes];
      // SAME_CLASS: [c0, c1);  OTHER_CLASSES: [0, c0) U [c1, cn)
      const int lo[2] = {mode == TIP_RANGE_SAME_CLASS ? c0 : 0, mode == TIP_RANGE_SAME_CLASS ? 0 : c1};
      const int hi[2] = {mode == TIP_RANGE_SAME_CLASS ? c1 : c0, mode == TIP_RANGE_SAME_CLASS ? 0 : cn};
      for (int rg = 0; rg < 2; rg++) {
        for (int base = lo[rg]; base < hi[rg]; base += kGroups) {
          const int j = base + grp;
          const bool ok = j < hi[rg];
          const T s = np_sumsq_g8<T>(x, t + (int64_t)(ok ? j : lo[rg]) * d, d, sub);
          if (ok && sub == 0) consider(best, Rn<T>::sqrt(s), t_gid ? t_gid[j] : j, j);
        }
      }
      if (threadIdx.x == 0 && stats) atomicAdd(stats + 0, 1ULL);
    }
    s_dist[threadIdx.x] = best.dist;
    s_gid[threadIdx.x] = best.gid;
    s_pos[threadIdx.x] = best.pos;
    __syncthreads();
    for (int o = kRerankThreads / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        Best<T> a{s_dist[threadIdx.x], s_gid[threadIdx.x], s_pos[threadIdx.x]};
        const int p2 = s_pos[threadIdx.x + o];
        if (p2 >= 0) {
          if (a.pos < 0) { a.dist = s_dist[threadIdx.x + o]; a.gid = s_gid[threadIdx.x + o]; a.pos = p2; }
          else consider(a, s_dist[threadIdx.x + o], s_gid[threadIdx.x + o], p2);
        }
        s_dist[threadIdx.x] = a.dist; s_gid[threadIdx.x] = a.gid; s_pos[threadIdx.x] = a.pos;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      out_dist[row] = s_pos[0] >= 0 ? s_dist[0] : (T)NAN;   // empty range -> NaN / -1
      out_pos[row] = s_pos[0];
    }
    __syncthreads();
  }
}

}  // namespace tip

using namespace tip;

extern "C" int tip_nn_rerank(const void* q, const void* t, int dtype, int64_t m, int64_t n, int64_t d,
                             const int32_t* cand_idx, const int32_t* cand_cnt, int32_t cap, const int32_t* q_class,
                             const int32_t* class_off, int32_t n_classes, int mode, const int32_t* t_gid,
                             void* out_dist, int32_t* out_pos, int64_t* stats, void* stream) {
  TIP_REQUIRE(q && t && out_dist && out_pos, "null pointer");
  TIP_REQUIRE(class_off && n_classes >= 1, "class offsets");
  TIP_REQUIRE(m >= 0 && n >= 0 && n < (1LL << 31) && d >= 1 && d < (1LL << 31), "shape");
  TIP_REQUIRE(mode == TIP_RANGE_SAME_CLASS || mode == TIP_RANGE_OTHER_CLASSES, "mode");
  TIP_REQUIRE(cand_cnt == nullptr || (cand_idx != nullptr && cap >= 1), "candidate buffers");
  if (m == 0) return TIP_OK;
  const int grid = (int)std::min<int64_t>(m, (int64_t)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t elem = dtype == TIP_F64 ? 8 : 4;
  const bool xs_ok = (size_t)d * elem <= 32 * 1024;
  const size_t xs_bytes = xs_ok ? (size_t)d * elem : 0;
  if (dtype == TIP_F32)
    rerank_kernel<float><<<grid, kRerankThreads, xs_bytes, st>>>((const float*)q, (const float*)t, m, n, (int)d, cand_idx,
                                                          cand_cnt, cap, q_class, class_off, n_classes, mode, t_gid,
                                                          (float*)out_dist, out_pos, (unsigned long long*)stats, xs_ok);
  else if (dtype == TIP_F64)
    rerank_kernel<double><<<grid, kRerankThreads, xs_bytes, st>>>((const double*)q, (const double*)t, m, n, (int)d,
                                                           cand_idx, cand_cnt, cap, q_class, class_off, n_classes,
                                                           mode, t_gid, (double*)out_dist, out_pos,
                                                           (unsigned long long*)stats, xs_ok);
  else
    TIP_REQUIRE(false, "dtype must be TIP_F32 or TIP_F64");
  TIP_LAUNCH_CHECK();
  return TIP_OK;
}
