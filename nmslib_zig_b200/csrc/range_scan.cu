// range_scan.cu -- range query of ONE query object against the whole shard, for seq_search indexes.
//
// Replaces SeqSearch<dist_t>::Search(RangeQuery*) (src/method/seqsearch.cc:108-141) +
// RangeQuery::CheckAndAddToResult (`distance <= radius`, src/rangequery.cc:58-65) as reached through
// nmslib_range_query_fill (nmslib_c.cpp:1051-1153): every object is visited in position order, the
// ones within the radius are reported in that order, truncated to the caller's capacity.
// One pass over the rows (HBM-bound: N * D * 4 bytes), exact fp32 distances with the reference's own
// formulas (the same code as the re-rank of scan_tc.cu), then an ordered collection by one block.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace nb200 {
namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// dist[i] = distance(row i, query); mode SCAN_L2 (sqrt when take_sqrt), SCAN_L1, SCAN_LINF, SCAN_COSINE,
// SCAN_ANGULAR, SCAN_NEGDOT.  (db_norm2 is unused: the cosine family recomputes all three sums per pair.)
__global__ void __launch_bounds__(256) range_dist_kernel(const float* __restrict__ db, const float* __restrict__ q,
                                                         const float* __restrict__ db_norm2, int n, int row_words,
                                                         int mode, int take_sqrt, float* __restrict__ dist) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rw4 = row_words >> 2;
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = blockIdx.x * (blockDim.x >> 5) + warp; row < n; row += warps) {
    if (mode == SCAN_SIFT) {  // byte rows of 128 (row_words = 32): exact int32 sum (x - y)^2, reported as a float
      const uchar4 x = __ldg(reinterpret_cast<const uchar4*>(db + (size_t)row * row_words) + lane);
      const uchar4 y = reinterpret_cast<const uchar4*>(q)[lane];
      const int d0 = (int)x.x - y.x, d1 = (int)x.y - y.y, d2 = (int)x.z - y.z, d3 = (int)x.w - y.w;
      int di = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      di = __reduce_add_sync(FULL, di);
      if (lane == 0) dist[row] = (float)di;
      continue;
    }
    const float4* x4 = reinterpret_cast<const float4*>(db + (size_t)row * row_words);
    float acc = 0.f, nxs = 0.f, nqs = 0.f;
    for (int e = lane; e < rw4; e += 32) {
      const float4 x = __ldg(x4 + e), y = q4[e];
      if (mode == SCAN_L2) {
        const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        acc = fmaf(d0, d0, acc);
        acc = fmaf(d1, d1, acc);
        acc = fmaf(d2, d2, acc);
        acc = fmaf(d3, d3, acc);
      } else if (mode == SCAN_L1) {
        acc += fabsf(x.x - y.x) + fabsf(x.y - y.y) + fabsf(x.z - y.z) + fabsf(x.w - y.w);
      } else if (mode == SCAN_LINF) {
        acc = fmaxf(acc, fmaxf(fmaxf(fabsf(x.x - y.x), fabsf(x.y - y.y)), fmaxf(fabsf(x.z - y.z), fabsf(x.w - y.w))));
      } else {
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
        acc = fmaf(x.z, y.z, acc);
        acc = fmaf(x.w, y.w, acc);
        if (mode == SCAN_COSINE || mode == SCAN_ANGULAR) {  // one summation pattern for all three sums (see scan_tc.cu)
          nxs = fmaf(x.x, x.x, nxs);
          nxs = fmaf(x.y, x.y, nxs);
          nxs = fmaf(x.z, x.z, nxs);
          nxs = fmaf(x.w, x.w, nxs);
          nqs = fmaf(y.x, y.x, nqs);
          nqs = fmaf(y.y, y.y, nqs);
          nqs = fmaf(y.z, y.z, nqs);
          nqs = fmaf(y.w, y.w, nqs);
        }
      }
    }
    if (mode == SCAN_LINF) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc = fmaxf(acc, __shfl_xor_sync(FULL, acc, o));
    } else {
      acc = warp_sum(acc);
      if (mode == SCAN_COSINE || mode == SCAN_ANGULAR) {
        nxs = warp_sum(nxs);
        nqs = warp_sum(nqs);
      }
    }
    if (lane == 0) {
      float d;
      if (mode == SCAN_L2) {
        d = take_sqrt ? sqrtf(acc) : acc;
      } else if (mode == SCAN_L1 || mode == SCAN_LINF) {
        d = acc;
      } else if (mode == SCAN_NEGDOT) {
        d = -acc;
      } else {
        const float eps = 2.0f * 1.17549435e-38f;
        float nsp = 0.f;
        if (!(nxs < eps || nqs < eps)) nsp = fmaxf(-1.f, fminf(1.f, acc / sqrtf(nxs) / sqrtf(nqs)));
        d = mode == SCAN_ANGULAR ? acosf(nsp) : fmaxf(0.f, 1.f - nsp);
      }
      dist[row] = d;
    }
  }
}

// rows with dist <= radius, in position order, at most `capacity` of them; *out_count = how many were written
__global__ void __launch_bounds__(1024) range_collect_kernel(const float* __restrict__ dist,
                                                             const int32_t* __restrict__ ext_ids, int n, float radius,
                                                             int capacity, int32_t* __restrict__ out_ids,
                                                             float* __restrict__ out_dists, int* __restrict__ out_count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    const float d = i < n ? dist[i] : 0.f;
    const bool hit = i < n && d <= radius;
    const unsigned m = __ballot_sync(FULL, hit);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    const int at = off + __popc(m & ((1u << lane) - 1));
    if (hit && at < capacity) {
      out_ids[at] = ext_ids[i];
      out_dists[at] = d;
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
    if (s_base >= capacity) break;  // (uniform: read after the barrier)
  }
  if (tid == 0) *out_count = min(s_base, capacity);
}

}  // namespace

cudaError_t launch_range_scan(const float* db, const float* query, const float* db_norm2, const int32_t* ext_ids, int n,
                              int row_words, int mode, int take_sqrt, float radius, int capacity, float* dist_tmp,
                              int32_t* out_ids, float* out_dists, int* out_count, cudaStream_t stream) {
  if (n <= 0 || capacity <= 0) return cudaErrorInvalidValue;
  if (row_words % 4) return cudaErrorInvalidValue;
  const int blocks = (int)std::min<long>(((long)n + 7) / 8, 148L * 8);
  range_dist_kernel<<<blocks, 256, 0, stream>>>(db, query, db_norm2, n, row_words, mode, take_sqrt, dist_tmp);
  range_collect_kernel<<<1, 1024, 0, stream>>>(dist_tmp, ext_ids, n, radius, capacity, out_ids, out_dists, out_count);
  return cudaGetLastError();
}

}  // namespace nb200
