// tc_peak.cu -- tcgen05 micro-benchmarks on one B200 ("TF32 and INT8 tensor peaks: measure with a
// tcgen05 micro-benchmark on the box and record beside the result").
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/tc_peak tools/tc_peak.cu
//   tools/tc_peak [seconds_sustained=4]  > profiles/r02_tc_peak.json
//
// Measures, with one CTA (or CTA pair) per SM issuing a back-to-back stream of MMAs on random operands that stay
// resident in shared / tensor memory (no memory traffic: this is the tensor pipe's ceiling, not a GEMM):
//   * kind::tf32  M128 N256 K8   A,B in shared memory (SS) / A in tensor memory (TS), cta_group::1
//   * kind::tf32  M256 N256 K8   cta_group::2 (SS)
//   * kind::i8    M128 N256 K32  SS / TS, cta_group::1;  M256 N256 K32 cta_group::2
//   burst (best of 10 launches of ~2 ms) and sustained (launches back to back for N seconds, clocks sampled by the
//   caller); TFLOP/s = 2 M N K per MMA.
//   * tcgen05.ld drain rate (bytes / clk / SM) with 4, 8 and 16 warps and the x32 / x64 / x128 shapes
//   * a correctness probe of kind::i8 (u8 x u8 -> s32, SS and TS operand layouts) against the CPU
// Prints one JSON object.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

namespace {
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_LOOP:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra W_DONE;\n\tbra "
      "W_LOOP;\n\tW_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int G>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  if (G == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int G>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  if (G == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int G>
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  if (G == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
enum Kind { K_TF32 = 0, K_I8 = 1 };
// D[tmem] (+)= A * B^T; A from shared memory (descriptor) or tensor memory (address)
template <int KIND, int G, bool TS>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == K_TF32) {
    if (TS)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                   "r"((uint32_t)a), "l"(b), "r"(idesc), "r"(acc)
                   : "memory");
    else if (G == 1)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a), "l"(b), "r"(idesc), "r"(acc)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a), "l"(b), "r"(idesc), "r"(acc)
                   : "memory");
  } else {
    if (TS)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                   "r"((uint32_t)a), "l"(b), "r"(idesc), "r"(acc)
                   : "memory");
    else if (G == 1)
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a), "l"(b), "r"(idesc), "r"(acc)
                   : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                   "l"(a), "l"(b), "r"(idesc), "r"(acc)
                   : "memory");
  }
}
// K-major operand tile, rows of 128 bytes, 128B swizzle, 8-row atoms of 1024 B (the layout a TMA box produces)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int kind, int m, int n) {
  // tf32: D = F32 (1 @4), A = B = TF32 (2 @7, 2 @10); i8: D = S32 (2 @4), A = B = unsigned 8-bit (0 @7, 0 @10)
  return (kind == K_TF32 ? ((1u << 4) | (2u << 7) | (2u << 10)) : (2u << 4)) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t swz_off(int row, int kbyte) {  // byte offset of (row, kbyte) in a swizzled tile
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kbyte >> 4) ^ (row & 7)) << 4) | (kbyte & 15)));
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int X>
__device__ __forceinline__ uint32_t tmem_ld_x(uint32_t taddr);
// ------------------------------------------------------------------------------------------------ MMA stream
// One CTA per SM (G == 2: one pair per two SMs), 160 threads: warps 0-3 fill the operands, warp 4 issues.
// `batches` x 16 MMAs, commits double buffered so the pipe never drains.
template <int KIND, int G, bool TS, int N>
__global__ void __launch_bounds__(416, 1) mma_stream_kernel(int batches, unsigned long long* clk_out, int drain) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int B_ROWS = N / G;  // each CTA of a pair holds its half of B
  __shared__ volatile int done_flag;
  if (threadIdx.x == 0) done_flag = 0;
  unsigned char* sA = smem;                 // 128 rows x 128 B
  unsigned char* sB = smem + 16384;         // B_ROWS x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + B_ROWS * 128);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = __shfl_sync(FULL, tid >> 5, 0), lane = tid & 31;
  const bool leader = G == 1 || cluster_ctarank() == 0;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  // random operands (power draw depends on the data: zeros would flatter the clocks)
  for (int i = tid; i < (16384 + B_ROWS * 128) / 4; i += blockDim.x) {
    uint32_t h = hash32(i * 2654435761u + blockIdx.x * 97u + 12345u);
    if (KIND == K_TF32) h = ((h & 0x007FE000u) | 0x3F800000u) ^ (h & 0x80000000u);  // +-[1,2), TF32-exact
    reinterpret_cast<uint32_t*>(smem)[i] = h;
  }
  fence_proxy_async();
  __syncthreads();
  if (G == 2) cluster_sync_all();
  if (warp == 4) tmem_alloc<G>(holder, 512);
  tc_fence_before();
  __syncthreads();
  if (G == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  if (TS && warp < 4) {  // A operand rows -> tensor memory columns [0, 32): thread = lane = row
    uint32_t v[8];
    for (int c = 0; c < 8; ++c) {
      for (int j = 0; j < 8; ++j) {
        uint32_t h = hash32((tid * 64 + c * 8 + j) * 40503u + 777u);
        if (KIND == K_TF32) h = ((h & 0x007FE000u) | 0x3F800000u) ^ (h & 0x80000000u);
        v[j] = h;
      }
      tmem_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && leader) {
    constexpr uint32_t idesc = make_idesc(KIND, 128 * G, N);
    const uint64_t da = make_smem_desc(smem_u32(sA)), db = make_smem_desc(smem_u32(sB));
    const uint32_t d0 = tmem_base + 256;  // accumulators: columns [256, 512)
    const long long t0 = clock64();
    for (int b = 0; b < batches; ++b) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int ks = j & 3;  // four K steps of 32 bytes inside the 128-byte rows
          if (drain & 256) {  // alternate between two accumulators and two A row blocks (what a 256-query CTA does)
            const int hh = j & 1, k2 = (j >> 1) & 3;
            umma<KIND, G, TS>(d0 + hh * N, TS ? (uint64_t)(tmem_base + hh * 32 + 8 * k2) : da + 2 * k2, db + 2 * k2, idesc,
                              (b | (j >> 1)) != 0);
          } else if (drain & 512) {  // two accumulators, but runs of 8 MMAs per accumulator
            const int hh = (j >> 3) & 1;
            umma<KIND, G, TS>(d0 + hh * N, TS ? (uint64_t)(tmem_base + hh * 32 + 8 * ks) : da + 2 * ks, db + 2 * ks, idesc,
                              (b | (j & 7)) != 0);
          } else {
            umma<KIND, G, TS>(d0, TS ? (uint64_t)(tmem_base + 8 * ks) : da + 2 * ks, db + 2 * ks, idesc, (b | j) != 0);
          }
        }
        tc_commit<G>(&bars[b & 1]);
      }
      __syncwarp();
      if (b > 0) mbar_wait(&bars[(b - 1) & 1], ((b - 1) >> 1) & 1);
    }
    mbar_wait(&bars[(batches - 1) & 1], ((batches - 1) >> 1) & 1);
    const long long t1 = clock64();
    if (lane == 0 && clk_out) clk_out[blockIdx.x] = (unsigned long long)(t1 - t0);
    done_flag = 1;
  } else if (warp >= 5 && warp < 5 + (drain & 255)) {
    // concurrent accumulator drains (what an epilogue does): warp w reads lane quarter w % 4, 32 columns at a time
    const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256;
    uint32_t acc = 0;
    int i = 0;
    while (!done_flag) {
      acc ^= tmem_ld_x<32>(trow + (uint32_t)((i * 32) & (N - 1) & ~31));
      ++i;
    }
    if (acc == 0x12345678u && clk_out) clk_out[0] = acc;
  } else if (warp == 4 && G == 2) {
    // the peer's barriers receive the multicast commits too; nothing to do but stay resident
    for (int b = 0; b < batches; ++b)
      if (b > 0) mbar_wait(&bars[(b - 1) & 1], ((b - 1) >> 1) & 1);
    mbar_wait(&bars[(batches - 1) & 1], ((batches - 1) >> 1) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (G == 2) cluster_sync_all();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<G>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ TMEM drain rate
template <int X>
__device__ __forceinline__ uint32_t tmem_ld_x(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t tmem_ld_x<32>(uint32_t taddr) {
  uint32_t v[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= v[i];
  return x;
}
template <>
__device__ __forceinline__ uint32_t tmem_ld_x<64>(uint32_t taddr) {
  uint32_t v[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, "
      "%38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, "
      "%60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) x ^= v[i];
  return x;
}
template <>
__device__ __forceinline__ uint32_t tmem_ld_x<16>(uint32_t taddr) {
  uint32_t v[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  tmem_ld_wait();
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x ^= v[i];
  return x;
}

// `warps` warps per CTA (4, 8 or 16; warp w reads lane quarter w % 4), each draining `iters` x X columns.
template <int X>
__global__ void __launch_bounds__(544, 1) tmem_drain_kernel(int warps, int iters, unsigned long long* clk_out, uint32_t* sink) {
  __shared__ uint32_t holder;
  const int tid = threadIdx.x, warp = __shfl_sync(FULL, tid >> 5, 0);
  if (warp == 16) tmem_alloc<1>(&holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = holder;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < warps) {
    const uint32_t trow = base + ((uint32_t)((warp & 3) * 32) << 16);
    const int grp = warp >> 2;  // warps of one SMSP read different column ranges
    t0 = clock64();
    for (int i = 0; i < iters; ++i) acc ^= tmem_ld_x<X>(trow + (uint32_t)(((i + grp * 5) * X) & 511 & ~(X - 1)));
    t1 = clock64();
  }
  __syncthreads();
  if (warp < warps && (tid & 31) == 0) atomicMax(clk_out + blockIdx.x, (unsigned long long)(t1 - t0));
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc<1>(base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ i8 probe
// D[128][64] = A[128][128] (u8) * B[64][128]^T (u8), K = 4 x 32, against the host; A from shared or tensor memory.
template <bool TS>
__global__ void __launch_bounds__(160, 1) i8_probe_kernel(const uint8_t* A, const uint8_t* B, int32_t* D) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sB = smem + 16384;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint32_t* holder = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = __shfl_sync(FULL, tid >> 5, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  for (int i = tid; i < 128 * 128; i += blockDim.x) sA[swz_off(i >> 7, i & 127)] = A[i];
  for (int i = tid; i < 64 * 128; i += blockDim.x) sB[swz_off(i >> 7, i & 127)] = B[i];
  fence_proxy_async();
  if (warp == 4) tmem_alloc<1>(holder, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  if (TS && warp < 4) {  // row `tid`: 128 bytes = 32 columns, byte k of the row in column k / 4, bits 8 (k % 4)
    for (int c = 0; c < 4; ++c) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = reinterpret_cast<const uint32_t*>(A + tid * 128)[c * 8 + j];
      tmem_st8(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 8, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  if (warp == 4) {
    constexpr uint32_t idesc = make_idesc(K_I8, 128, 64);
    const uint64_t da = make_smem_desc(smem_u32(sA)), db = make_smem_desc(smem_u32(sB));
    if (elect_one()) {
      for (int ks = 0; ks < 4; ++ks)
        umma<K_I8, 1, TS>(tmem_base + 256, TS ? (uint64_t)(tmem_base + 8 * ks) : da + 2 * ks, db + 2 * ks, idesc, ks != 0);
      tc_commit<1>(bar);
    }
    __syncwarp();
  }
  if (warp < 4) {
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int c = 0; c < 4; ++c) {
      uint32_t v[16];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(tmem_base + ((uint32_t)(warp * 32) << 16) + 256 + c * 16)
          : "memory");
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) D[tid * 64 + c * 16 + j] = (int32_t)v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

struct Rate {
  double burst = 0, sustained = 0, clk_per_mma = 0;
};

template <int KIND, int G, bool TS, int N = 256>
Rate run_stream(int sms, double seconds, unsigned long long* d_clk, int drain = 0) {
  auto kern = mma_stream_kernel<KIND, G, TS, N>;
  const int smem = 16384 + (N / G) * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = G == 2 ? (sms / 2) * 2 : sms;
  const int K = KIND == K_TF32 ? 8 : 32;
  const int batches = (KIND == K_TF32 ? 1500 : 6000) * (256 / N);  // ~2 ms per launch
  const double flop_per_launch = 2.0 * (128.0 * G) * (double)N * K * 16.0 * batches * (grid / G);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(416);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = G;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  Rate r;
  for (int i = 0; i < 3; ++i) CK(cudaLaunchKernelEx(&cfg, kern, batches, d_clk, drain));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 10; ++i) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kern, batches, d_clk, drain));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    r.burst = std::max(r.burst, flop_per_launch / (ms * 1e-3) / 1e12);
  }
  std::vector<unsigned long long> clk(grid);
  CK(cudaMemcpy(clk.data(), d_clk, grid * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0;
  for (int i = 0; i < grid; i += G) mx = std::max(mx, clk[i]);
  r.clk_per_mma = (double)mx / (16.0 * batches);
  if (seconds > 0) {
    int launches = 0;
    CK(cudaEventRecord(e0));
    auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
      for (int i = 0; i < 50; ++i) CK(cudaLaunchKernelEx(&cfg, kern, batches, d_clk, drain));
      launches += 50;
      CK(cudaStreamSynchronize(0));
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    r.sustained = flop_per_launch * launches / (ms * 1e-3) / 1e12;
  }
  return r;
}

template <int X>
double run_drain(int sms, int warps, unsigned long long* d_clk, uint32_t* d_sink) {
  const int iters = 20000;
  CK(cudaMemset(d_clk, 0, sms * 8));
  tmem_drain_kernel<X><<<sms, 544>>>(warps, iters, d_clk, d_sink);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(d_clk, 0, sms * 8));
  tmem_drain_kernel<X><<<sms, 544>>>(warps, iters, d_clk, d_sink);
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> clk(sms);
  CK(cudaMemcpy(clk.data(), d_clk, sms * 8, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0;
  for (auto c : clk) mx = std::max(mx, c);
  return (double)warps * iters * 32.0 * X * 4.0 / (double)mx;  // bytes per clock per SM
}

template <bool TS>
int run_probe() {
  std::vector<uint8_t> A(128 * 128), B(64 * 128);
  uint32_t s = 12345;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return (uint8_t)(s >> 24);
  };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  for (int i = 0; i < 128; ++i) A[5 * 128 + i] = 255, B[7 * 128 + i] = 255;  // the extreme: 128 * 255^2
  uint8_t *dA, *dB;
  int32_t* dD;
  CK(cudaMalloc(&dA, A.size()));
  CK(cudaMalloc(&dB, B.size()));
  CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  const int smem = 16384 + 64 * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(i8_probe_kernel<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  i8_probe_kernel<TS><<<1, 160, smem>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> D(128 * 64);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      int32_t ref = 0;
      for (int k = 0; k < 128; ++k) ref += (int32_t)A[m * 128 + k] * (int32_t)B[n * 128 + k];
      if (ref != D[m * 64 + n]) {
        if (bad < 4) fprintf(stderr, "i8 probe (TS=%d) mismatch at (%d,%d): got %d want %d\n", (int)TS, m, n, D[m * 64 + n], ref);
        ++bad;
      }
    }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return bad;
}

}  // namespace

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 4.0;
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  unsigned long long* d_clk;
  uint32_t* d_sink;
  CK(cudaMalloc(&d_clk, 8 * 1024));
  CK(cudaMalloc(&d_sink, 64));
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);

  const int bad_ss = run_probe<false>();
  const int bad_ts = run_probe<true>();

  printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_max_mhz\": %.0f, \"sustained_seconds\": %.1f,\n", prop.name, sms,
         clock_khz / 1e3, seconds);
  printf(" \"i8_probe_mismatches\": {\"ss\": %d, \"ts\": %d},\n", bad_ss, bad_ts);
  auto emit = [&](const char* name, const Rate& r, const char* unit, bool last = false) {
    printf(" \"%s\": {\"burst\": %.1f, \"sustained\": %.1f, \"unit\": \"%s\", \"clk_per_mma\": %.1f}%s\n", name, r.burst,
           r.sustained, unit, r.clk_per_mma, last ? "" : ",");
    fflush(stdout);
  };
  emit("tf32_m128n256k8_ss_cta1", run_stream<K_TF32, 1, false>(sms, 0, d_clk), "TFLOP/s");
  emit("tf32_m128n256k8_ts_cta1", run_stream<K_TF32, 1, true>(sms, seconds, d_clk), "TFLOP/s");
  emit("tf32_m256n256k8_ss_cta2", run_stream<K_TF32, 2, false>(sms, seconds, d_clk), "TFLOP/s");
  emit("tf32_m128n128k8_ts_cta1", run_stream<K_TF32, 1, true, 128>(sms, 0, d_clk), "TFLOP/s");
  emit("tf32_m128n64k8_ts_cta1", run_stream<K_TF32, 1, true, 64>(sms, 0, d_clk), "TFLOP/s");
  emit("tf32_m128n64k8_ss_cta1", run_stream<K_TF32, 1, false, 64>(sms, 0, d_clk), "TFLOP/s");
  emit("tf32_m128n64k8_ts_cta1_drain8", run_stream<K_TF32, 1, true, 64>(sms, 0, d_clk, 8), "TFLOP/s");
  emit("tf32_m128n64k8_ts_cta1_alt2acc", run_stream<K_TF32, 1, true, 64>(sms, 0, d_clk, 256), "TFLOP/s");
  emit("tf32_m128n64k8_ts_cta1_alt2acc_drain8", run_stream<K_TF32, 1, true, 64>(sms, 0, d_clk, 256 + 8), "TFLOP/s");
  emit("tf32_m128n64k8_ts_cta1_runs8", run_stream<K_TF32, 1, true, 64>(sms, 0, d_clk, 512), "TFLOP/s");
  emit("i8_m128n64k32_ts_cta1_alt2acc", run_stream<K_I8, 1, true, 64>(sms, 0, d_clk, 256), "TOP/s");
  emit("tf32_m128n256k8_ts_cta1_drain8", run_stream<K_TF32, 1, true, 256>(sms, 0, d_clk, 8), "TFLOP/s");
  emit("i8_m128n64k32_ts_cta1", run_stream<K_I8, 1, true, 64>(sms, 0, d_clk), "TOP/s");
  emit("i8_m128n128k32_ts_cta1", run_stream<K_I8, 1, true, 128>(sms, 0, d_clk), "TOP/s");
  emit("i8_m128n128k32_ts_cta1_drain8", run_stream<K_I8, 1, true, 128>(sms, 0, d_clk, 8), "TOP/s");
  emit("i8_m128n256k32_ss_cta1", run_stream<K_I8, 1, false>(sms, 0, d_clk), "TOP/s");
  emit("i8_m128n256k32_ts_cta1", run_stream<K_I8, 1, true>(sms, seconds, d_clk), "TOP/s");
  emit("i8_m256n256k32_ss_cta2", run_stream<K_I8, 2, false>(sms, seconds, d_clk), "TOP/s");
  printf(" \"tmem_ld_bytes_per_clk_per_sm\": {");
  const int ws[3] = {4, 8, 16};
  for (int i = 0; i < 3; ++i) {
    printf("\"x16_w%d\": %.1f, ", ws[i], run_drain<16>(sms, ws[i], d_clk, d_sink));
    printf("\"x32_w%d\": %.1f, ", ws[i], run_drain<32>(sms, ws[i], d_clk, d_sink));
    printf("\"x64_w%d\": %.1f%s", ws[i], run_drain<64>(sms, ws[i], d_clk, d_sink), i == 2 ? "" : ", ");
  }
  printf("}\n}\n");
  return (bad_ss || bad_ts) ? 3 : 0;
}
