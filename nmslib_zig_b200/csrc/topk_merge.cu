// topk_merge.cu -- K4: k-way merge of sorted (distance, position) key lists + result
// finalisation.  Replaces the per-thread queue merge of SeqSearch (seqsearch.cc:151-175),
// and extract_knn_results (nmslib_c.cpp:293-328: pop, reverse, cast to float, external id).
// It is used three ways: (1) to combine the per-split lists of one scan launch, (2) to
// combine the per-GPU lists after the NVLink all-gather, (3) with lists == 1
// as the plain "decode keys -> (id, float distance)" finaliser.
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

namespace nb200 {
namespace {

constexpr int MAX_ITEMS = 8192;  // keys per query that one CTA sorts in shared memory

__device__ __forceinline__ float decode_dist(uint64_t key, int finalize) {
  const uint32_t hi = (uint32_t)(key >> 32);
  if (finalize == FIN_INT) return (float)i32_from_ordered(hi);   // nmslib_c.cpp:317 int -> float
  const float v = f32_from_ordered(hi);
  return finalize == FIN_SQRT ? sqrtf(v) : v;                    // distcomp_lp.cc:368-371
}

__global__ void merge_topk_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ ids_in,
                                  int lists, size_t list_stride, size_t query_stride, int nq, int k,
                                  int items_pow2, int finalize, const int32_t* __restrict__ ext_ids,
                                  uint32_t pos_base, uint64_t* __restrict__ out_keys,
                                  int32_t* __restrict__ out_ids, float* __restrict__ out_dists,
                                  int32_t* __restrict__ out_counts, const int* __restrict__ d_nq) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // (query count decided on the device: launched with a small grid that strides over the -- normally zero -- queries,
  // so the launch that finds nothing to do costs ~2 us instead of scheduling nq blocks that leave at once)
  if (d_nq) nq = min(nq, *d_nq);
  uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
  int32_t* sid = reinterpret_cast<int32_t*>(sk + items_pow2);
  const int items = lists * k;
  for (int q = blockIdx.x; q < nq; q += gridDim.x) {

  for (int t = threadIdx.x; t < items_pow2; t += blockDim.x) {
    uint64_t key = KEY_MAX;
    int32_t id = -1;
    if (t < items) {
      const int l = t / k, e = t - l * k;
      const size_t off = (size_t)l * list_stride + (size_t)q * query_stride + e;
      key = keys[off];
      if (ids_in) id = ids_in[off];
    }
    sk[t] = key;
    sid[t] = id;
  }
  __syncthreads();

  if (lists > 1) {  // bitonic sort, ascending
    for (int size = 2; size <= items_pow2; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < items_pow2 / 2; t += blockDim.x) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool up = (lo & size) == 0;
          const uint64_t a = sk[lo], b = sk[hi];
          if ((a > b) == up) {
            sk[lo] = b;
            sk[hi] = a;
            const int32_t ia = sid[lo];
            sid[lo] = sid[hi];
            sid[hi] = ia;
          }
        }
        __syncthreads();
      }
    }
  }

  int local_cnt = 0;
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    const uint64_t key = e < items_pow2 ? sk[e] : KEY_MAX;
    const size_t o = (size_t)q * k + e;
    int32_t id = -1;
    float d = __int_as_float(0x7F800000);
    if (key != KEY_MAX) {
      ++local_cnt;
      d = decode_dist(key, finalize);
      if (ids_in) id = sid[e];
      else {
        const uint32_t pos = (uint32_t)key;
        id = ext_ids ? ext_ids[pos - pos_base] : (int32_t)pos;
      }
    }
    if (out_keys) out_keys[o] = key;
    if (out_ids) out_ids[o] = id;
    if (out_dists) out_dists[o] = d;
  }
  if (out_counts) {
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (local_cnt) atomicAdd(&total, local_cnt);
    __syncthreads();
    if (threadIdx.x == 0) out_counts[q] = total;
  }
  __syncthreads();  // (the next query of this block reuses the shared arrays)
  }
}

}  // namespace

namespace {
// d_count (optional): the number of rows is decided on the device; rows [count, count rounded up to `pad`) are
// zero-filled so that the consumer's last query block reads defined padding, blocks past that leave at once
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ idx, int count,
                                   int row_vec, uint4* __restrict__ dst, const int* __restrict__ d_count, int pad) {
  const int c = d_count ? min(count, *d_count) : count;
  const int lim = d_count ? (c + pad - 1) / pad * pad : count;  // (a small grid strides over the rows)
  for (int r = blockIdx.x; r < lim; r += gridDim.x) {
    uint4* d = dst + (size_t)r * row_vec;
    if (r >= c) {
      for (int j = threadIdx.x; j < row_vec; j += blockDim.x) d[j] = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const uint4* s = src + (size_t)idx[r] * row_vec;
    for (int j = threadIdx.x; j < row_vec; j += blockDim.x) d[j] = s[j];
  }
}
__global__ void scatter_keys_kernel(const uint64_t* __restrict__ src, const int* __restrict__ idx, int count, int k,
                                    uint64_t* __restrict__ dst, const int* __restrict__ d_count, int f2i) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (d_count) count = min(count, *d_count);
  if (i >= count * k) return;
  if (f2i) {
    const int r = i / k, e = i - r * k;
    const uint64_t key = src[i];
    dst[(size_t)idx[r] * k + e] = key == KEY_MAX ? key : make_key(i32_ordered((int)f32_from_ordered((uint32_t)(key >> 32))), (uint32_t)key);
    return;
  }
  const int r = i / k, e = i - r * k;
  dst[(size_t)idx[r] * k + e] = src[i];
}
}  // namespace

namespace {
__global__ void widen_u8_kernel(const uint8_t* __restrict__ src, size_t rows, int dim, int row_words,
                                float* __restrict__ dst) {
  const size_t total = rows * (size_t)(dim >> 2);  // one uchar4 per thread step
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / (dim >> 2), c = (i % (dim >> 2)) * 4;
    const uchar4 v = *reinterpret_cast<const uchar4*>(src + r * dim + c);
    *reinterpret_cast<float4*>(dst + r * row_words + c) = make_float4(v.x, v.y, v.z, v.w);
  }
}
}  // namespace

cudaError_t launch_widen_u8(const uint8_t* src, size_t rows, int dim, int row_words, float* dst, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  if (dim % 4 || row_words % 4) return cudaErrorInvalidValue;
  const size_t total = rows * (size_t)(dim >> 2);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  widen_u8_kernel<<<blocks, 256, 0, stream>>>(src, rows, dim, row_words, dst);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const uint32_t* src, const int* idx, int count, int row_words, uint32_t* dst,
                               cudaStream_t stream, const int* d_count, int pad) {
  if (count <= 0) return cudaSuccess;
  const int blocks = std::min(d_count ? (count + pad - 1) / pad * pad : count, 1184);
  gather_rows_kernel<<<blocks, 64, 0, stream>>>(reinterpret_cast<const uint4*>(src), idx, count, row_words / 4,
                                                reinterpret_cast<uint4*>(dst), d_count, pad);
  return cudaGetLastError();
}
cudaError_t launch_scatter_keys(const uint64_t* src, const int* idx, int count, int k, uint64_t* dst,
                                cudaStream_t stream, const int* d_count, int f2i) {
  if (count <= 0) return cudaSuccess;
  const int total = count * k;
  scatter_keys_kernel<<<(total + 255) / 256, 256, 0, stream>>>(src, idx, count, k, dst, d_count, f2i);
  return cudaGetLastError();
}

int merge_topk_max_items() { return MAX_ITEMS; }

cudaError_t launch_merge_topk(const uint64_t* keys, const int32_t* ids_in, int lists, size_t list_stride,
                              size_t query_stride, int nq, int k, int finalize, const int32_t* ext_ids,
                              uint32_t pos_base, uint64_t* out_keys, int32_t* out_ids, float* out_dists,
                              int32_t* out_counts, cudaStream_t stream, const int* d_nq) {
  if (nq <= 0) return cudaSuccess;
  const int items = lists * k;
  if (items > MAX_ITEMS) return cudaErrorInvalidValue;
  int p2 = 1;
  while (p2 < items) p2 <<= 1;
  int threads = p2 / 2;
  if (threads < 32) threads = 32;
  if (threads > 256) threads = 256;
  const size_t smem = (size_t)p2 * 12 + 16;
  cudaError_t e =
      cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAX_ITEMS * 12 + 16));
  if (e != cudaSuccess) return e;
  const int blocks = d_nq ? std::min(nq, 1184) : nq;
  merge_topk_kernel<<<blocks, threads, smem, stream>>>(keys, ids_in, lists, list_stride, query_stride, nq, k, p2,
                                                   finalize, ext_ids, pos_base, out_keys, out_ids, out_dists,
                                                   out_counts, d_nq);
  return cudaGetLastError();
}

}  // namespace nb200
