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
              int32_t* __restrict__ out_pos, unsigned long long* __restrict__ stats) {
  __shared__ T s_dist[kRerankThreads];
  __shared__ int s_gid[kRerankThreads];
  __shared__ int s_pos[kRerankThreads];
  for (int64_t row = blockIdx.x; row < m; row += gridDim.x) {
    const T* x = q + row * (int64_t)d;
    Best<T> best;
    best.dist = Rn<T>::inf();
    best.gid = 0x7fffffff;
    best.pos = -1;
    const int cnt = cand_cnt ? cand_cnt[row] : 0;
    const int cls = q_class ? q_class[row] : 0;
    const bool listed = cand_cnt != nullptr && cnt >= 1 && cnt <= cap;
    if (listed) {
      for (int k = threadIdx.x; k < cnt; k += kRerankThreads) {
        const int j = cand_idx[row * (int64_t)cap + k];
        if (j < 0 || j >= n) continue;
        const T s = np_sumsq<T>(x, t + (int64_t)j * d, d);
        consider(best, Rn<T>::sqrt(s), t_gid ? t_gid[j] : j, j);
      }
      if (threadIdx.x == 0 && stats) atomicAdd(stats + 1, (unsigned long long)cnt);
    } else if (cls >= 0 && cls < n_classes) {
      const int c0 = class_off[cls], c1 = class_off[cls + 1], cn = class_off[n_classes];
      // SAME_CLASS: [c0, c1);  OTHER_CLASSES: [0, c0) U [c1, cn)
      const int lo0 = mode == TIP_RANGE_SAME_CLASS ? c0 : 0;
      const int hi0 = mode == TIP_RANGE_SAME_CLASS ? c1 : c0;
      const int lo1 = mode == TIP_RANGE_SAME_CLASS ? 0 : c1;
      const int hi1 = mode == TIP_RANGE_SAME_CLASS ? 0 : cn;
      for (int j = lo0 + threadIdx.x; j < hi0; j += kRerankThreads) {
        const T s = np_sumsq<T>(x, t + (int64_t)j * d, d);
        consider(best, Rn<T>::sqrt(s), t_gid ? t_gid[j] : j, j);
      }
      for (int j = lo1 + threadIdx.x; j < hi1; j += kRerankThreads) {
        const T s = np_sumsq<T>(x, t + (int64_t)j * d, d);
        consider(best, Rn<T>::sqrt(s), t_gid ? t_gid[j] : j, j);
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
  if (dtype == TIP_F32)
    rerank_kernel<float><<<grid, kRerankThreads, 0, st>>>((const float*)q, (const float*)t, m, n, (int)d, cand_idx,
                                                          cand_cnt, cap, q_class, class_off, n_classes, mode, t_gid,
                                                          (float*)out_dist, out_pos, (unsigned long long*)stats);
  else if (dtype == TIP_F64)
    rerank_kernel<double><<<grid, kRerankThreads, 0, st>>>((const double*)q, (const double*)t, m, n, (int)d,
                                                           cand_idx, cand_cnt, cap, q_class, class_off, n_classes,
                                                           mode, t_gid, (double*)out_dist, out_pos,
                                                           (unsigned long long*)stats);
  else
    TIP_REQUIRE(false, "dtype must be TIP_F32 or TIP_F64");
  TIP_LAUNCH_CHECK();
  return TIP_OK;
}
