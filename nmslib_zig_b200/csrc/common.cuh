// common.cuh -- shared device/host helpers for the nmslib_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nb200 {

// Distance spaces served by the engine.
enum Space : int {
  SPACE_L2 = 0,        // seq_search: sqrtf(sum (x-y)^2)  (space_lp.h:57-58)
  SPACE_L2SQR = 1,     // sum (x-y)^2                     (distcomp_lp.cc:304-365)
  SPACE_COSINE = 2,    // max(0, 1 - clamp(x.y/|x|/|y|))  (distcomp_scalar.cc:84-168, 268-271)
  SPACE_NEGDOT = 3,    // -x.y                            (space_scalar.cc:60-68)
  SPACE_L2SQR_SIFT = 4,// int32 n1 + n2 - 2 x.y           (distcomp_l2sqr_sift.cc:41-151)
  SPACE_L1 = 5,        // sum |x-y|                       (distcomp_lp.cc:190-251)      seq_search only
  SPACE_LINF = 6,      // max |x-y|                       (distcomp_lp.cc:77-139)       seq_search only
  SPACE_ANGULAR = 7    // acos(clamp(x.y/|x|/|y|))        (distcomp_scalar.cc:254-258)  seq_search only
};

// How the value stored in the upper half of a key becomes the reported float.
enum Finalize : int { FIN_FLOAT = 0, FIN_SQRT = 1, FIN_INT = 2 };

// ---- result keys ---------------------------------------------------------
// A candidate is the 64-bit key  ordered(distance) << 32 | position.  Unsigned
// comparison of keys is the lexicographic (distance, position) order, i.e. exactly
// what the reference's KNNQueue produces when objects are visited in position order
// (knnqueue.h:55-64, 73-74): the k smallest keys are the answer.
static constexpr uint64_t KEY_MAX = 0xFFFFFFFFFFFFFFFFull;

__host__ __device__ __forceinline__ uint32_t f32_ordered(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f + 0.0f);  // -0 -> +0
#else
  union { float f; uint32_t u; } c; c.f = f + 0.0f; uint32_t b = c.u;
#endif
  if ((b & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;  // NaN sorts last, never selected
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_from_ordered(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint32_t i32_ordered(int32_t v) {
  return (uint32_t)v ^ 0x80000000u;
}
__host__ __device__ __forceinline__ int32_t i32_from_ordered(uint32_t u) {
  return (int32_t)(u ^ 0x80000000u);
}
__host__ __device__ __forceinline__ uint64_t make_key(uint32_t ordered, uint32_t pos) {
  return ((uint64_t)ordered << 32) | pos;
}

// ---- cp.async (LDGSTS) -----------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__host__ __device__ __forceinline__ size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

}  // namespace nb200
