// exchange.cu -- the cross-shard step of the row-sharded search over NVLink PEER MEMORY, inside the
// library: no collective call per step.
//
// The reference combines per-thread chunk results inside SeqSearch::Search itself (seqsearch.cc:151-175: contiguous
// chunks, a top-k queue per chunk, one merge).  Here a chunk is a GPU.  Every rank (one per GPU; ranks are processes
// under torchrun / MPI, or the devices of one process under a ShardGroup) owns a WINDOW in its own HBM:
//
//     flags[w]            uint32, slot p written by peer p: "my lists of step <value> are complete"
//     keys[2][q][k]       uint64 (ordered distance << 32 | global position), double buffered by step parity
//     ids [2][q][k]       int32 external ids of the same entries
//
// and maps the windows of all peers (cudaIpcOpenMemHandle across processes, peer access inside one process).
// Per query batch, on the engine's stream:
//     publish_kernel   sorted local top-k keys -> own window (+ external ids), then ONE release-store per peer of
//                      the step number into that peer's flags[my rank]   (remote NVLink stores, fire and forget)
//     merge_kernel     spins (acquire loads of its OWN flags, local HBM) until every peer has published this step,
//                      then reads the peers' key / id lists straight over NVLink, merges k-way by key in shared
//                      memory and finalises (sqrt for l2, int -> float for l2sqr_sift)
// A window slot of parity s is rewritten at step s + 2, which its owner can only reach after its merge of step
// s + 1 saw every peer's flag of s + 1 -- written after that peer's merge of step s in stream order -- so no reader
// of step s is still active: two buffers suffice.  The kernels that wait run on DIFFERENT GPUs than the kernels
// they wait for; a wait that sees no progress for 10 s sets the window's error word instead of hanging.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>

#include "common.cuh"
#include "engine.h"
#include "kernels.h"

namespace nb200 {
namespace {

constexpr uint64_t kBlobMagic = 0x4E42323030584348ull;  // "NB200XCH"
constexpr int kHdrBytes = 1024;                          // flags[64] | ticket | err | pad
constexpr int MERGE_ITEMS = 8192;

struct Blob {  // NMSLIB_B200_SHARD_BLOB_BYTES = 256
  uint64_t magic, pid, ptr, bytes, cap_q, cap_k;
  int32_t device, pad;
  cudaIpcMemHandle_t handle;  // 64 bytes
};
static_assert(sizeof(Blob) <= 256, "blob must fit NMSLIB_B200_SHARD_BLOB_BYTES");

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int32_t ld_relaxed_sys_s32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct PeerPtrs {
  unsigned char* win[kMaxShardWorld];
};

__global__ void publish_kernel(const uint64_t* __restrict__ local_keys, const int32_t* __restrict__ ext_ids,
                               uint32_t pos_base, size_t items, unsigned char* my_win, size_t key_off, size_t id_off,
                               PeerPtrs peers, int rank, int world, uint32_t step) {
  uint64_t* wk = reinterpret_cast<uint64_t*>(my_win + key_off);
  int32_t* wi = reinterpret_cast<int32_t*>(my_win + id_off);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (size_t)gridDim.x * blockDim.x) {
    const uint64_t key = local_keys[i];
    wk[i] = key;
    wi[i] = key == KEY_MAX ? -1 : (ext_ids ? ext_ids[(uint32_t)key - pos_base] : (int32_t)(uint32_t)key);
  }
  // the last block to finish tells every peer (one remote store each) that this rank's lists of `step` are complete
  __threadfence();
  __shared__ unsigned last;
  unsigned* ticket = reinterpret_cast<unsigned*>(my_win + 256);
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) *ticket = 0u;
  __threadfence_system();
  if (threadIdx.x < world)
    st_release_sys(reinterpret_cast<uint32_t*>(peers.win[threadIdx.x]) + rank, step);
}

__device__ __forceinline__ float decode_dist_x(uint64_t key, int finalize) {
  const uint32_t hi = (uint32_t)(key >> 32);
  if (finalize == FIN_INT) return (float)i32_from_ordered(hi);
  const float v = f32_from_ordered(hi);
  return finalize == FIN_SQRT ? sqrtf(v) : v;
}

// one block per query of [q_begin, q_begin + gridDim.x); lists = world
__global__ void merge_peers_kernel(PeerPtrs peers, unsigned char* my_win, size_t key_off, size_t id_off, int world,
                                   uint32_t step, int q_begin, int k, int items_pow2, int finalize,
                                   uint64_t* __restrict__ out_keys, int32_t* __restrict__ out_ids,
                                   float* __restrict__ out_dists, int32_t* __restrict__ out_counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
  int32_t* sid = reinterpret_cast<int32_t*>(sk + items_pow2);
  const int q = q_begin + blockIdx.x;
  if (threadIdx.x < world) {
    const uint32_t* flag = reinterpret_cast<const uint32_t*>(my_win) + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    // (steps are compared as a signed distance: the counter may wrap)
    while ((int32_t)(ld_acquire_sys(flag) - step) < 0) {
      if (globaltimer_ns() - t0 > 10000000000ull) {  // a peer that never arrives: report, do not hang
        atomicExch(reinterpret_cast<unsigned*>(my_win + 260), 1u);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  const int items = world * k;
  for (int t = threadIdx.x; t < items_pow2; t += blockDim.x) {
    uint64_t key = KEY_MAX;
    int32_t id = -1;
    if (t < items) {
      const int l = t / k, e = t - l * k;
      const size_t off = (size_t)q * k + e;
      key = ld_relaxed_sys_u64(reinterpret_cast<const uint64_t*>(peers.win[l] + key_off) + off);
      id = ld_relaxed_sys_s32(reinterpret_cast<const int32_t*>(peers.win[l] + id_off) + off);
    }
    sk[t] = key;
    sid[t] = id;
  }
  __syncthreads();
  for (int size = 2; size <= items_pow2; size <<= 1) {  // bitonic sort, ascending by (distance, global position)
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
  int local_cnt = 0;
  for (int e = threadIdx.x; e < k; e += blockDim.x) {
    const uint64_t key = e < items_pow2 ? sk[e] : KEY_MAX;
    const size_t o = (size_t)q * k + e;
    const bool hit = key != KEY_MAX;
    local_cnt += hit ? 1 : 0;
    if (out_keys) out_keys[o] = key;
    if (out_ids) out_ids[o] = hit ? sid[e] : -1;
    if (out_dists) out_dists[o] = hit ? decode_dist_x(key, finalize) : __int_as_float(0x7F800000);
  }
  if (out_counts) {
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (local_cnt) atomicAdd(&total, local_cnt);
    __syncthreads();
    if (threadIdx.x == 0) out_counts[q] = total;
  }
}

}  // namespace

struct PeerExchange {
  int device = 0, rank = -1, world = 0;
  size_t cap_q = 0, cap_k = 0, bytes = 0;
  unsigned char* win = nullptr;          // this rank's window (cudaMalloc)
  unsigned char* peer[kMaxShardWorld] = {};
  bool opened[kMaxShardWorld] = {};      // mapped with cudaIpcOpenMemHandle (to be closed)
  uint32_t step = 0;
  size_t key_off(int parity) const { return kHdrBytes + (size_t)parity * cap_q * cap_k * 8; }
  size_t id_off(int parity) const { return kHdrBytes + 2 * cap_q * cap_k * 8 + (size_t)parity * cap_q * cap_k * 4; }
};

void xch_destroy(PeerExchange* x) {
  if (!x) return;
  cudaSetDevice(x->device);
  for (int p = 0; p < x->world; ++p)
    if (x->opened[p] && x->peer[p]) cudaIpcCloseMemHandle(x->peer[p]);
  if (x->win) cudaFree(x->win);
  cudaGetLastError();
  delete x;
}

Status xch_export(PeerExchange** out, int device, size_t max_q, size_t max_k, void* blob256) {
  if (!out || !blob256 || max_q == 0 || max_k == 0) return Status::Err(2, "invalid exchange window request");
  if (*out) {
    xch_destroy(*out);
    *out = nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) return Status::Err(9, "cudaSetDevice failed");
  PeerExchange* x = new PeerExchange();
  x->device = device;
  x->cap_q = max_q;
  x->cap_k = max_k;
  x->bytes = round_up(kHdrBytes + 2 * max_q * max_k * 12, 1 << 21);  // (IPC exports whole allocations: keep it 2 MB-granular)
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->win), x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->win, 0, kHdrBytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    delete x;
    return Status::Err(e == cudaErrorMemoryAllocation ? 3 : 9, std::string("exchange window: ") + cudaGetErrorString(e));
  }
  Blob b;
  memset(&b, 0, sizeof(b));
  b.magic = kBlobMagic;
  b.pid = (uint64_t)getpid();
  b.ptr = (uint64_t)(uintptr_t)x->win;
  b.bytes = x->bytes;
  b.cap_q = max_q;
  b.cap_k = max_k;
  b.device = device;
  if (cudaIpcGetMemHandle(&b.handle, x->win) != cudaSuccess) {
    cudaGetLastError();  // (no IPC in this environment: ranks of one process can still connect through the raw pointer)
    memset(&b.handle, 0, sizeof(b.handle));
  }
  memset(blob256, 0, 256);
  memcpy(blob256, &b, sizeof(b));
  *out = x;
  return Status::OK();
}

Status xch_connect(PeerExchange* x, int rank, int world, const void* blobs) {
  if (!x || !x->win) return Status::Err(2, "export the exchange window first");
  if (world < 1 || world > kMaxShardWorld || rank < 0 || rank >= world || !blobs)
    return Status::Err(2, "invalid rank / world size for the shard exchange");
  if (cudaSetDevice(x->device) != cudaSuccess) return Status::Err(9, "cudaSetDevice failed");
  const uint64_t me = (uint64_t)getpid();
  for (int p = 0; p < world; ++p) {
    Blob b;
    memcpy(&b, static_cast<const unsigned char*>(blobs) + (size_t)p * 256, sizeof(b));
    if (b.magic != kBlobMagic) return Status::Err(2, "shard exchange blob " + std::to_string(p) + " is not a blob");
    if (b.cap_q != x->cap_q || b.cap_k != x->cap_k)
      return Status::Err(2, "all ranks must export windows of the same shape");
    if (p == rank) {
      if (b.ptr != (uint64_t)(uintptr_t)x->win || b.pid != me) return Status::Err(2, "blob of this rank is not its own");
      x->peer[p] = x->win;
    } else if (b.pid == me) {  // a device of this process: peer access, raw pointer
      if (b.device != x->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, x->device, b.device);
        if (!can) return Status::Err(9, "no peer access between devices " + std::to_string(x->device) + " and " + std::to_string(b.device));
        cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return Status::Err(9, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        cudaGetLastError();
      }
      x->peer[p] = reinterpret_cast<unsigned char*>((uintptr_t)b.ptr);
    } else {  // another process: CUDA IPC
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return Status::Err(9, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(p) + "): " + cudaGetErrorString(e));
      }
      x->peer[p] = static_cast<unsigned char*>(ptr);
      x->opened[p] = true;
    }
  }
  x->rank = rank;
  x->world = world;
  x->step = 0;
  return Status::OK();
}

bool xch_connected(const PeerExchange* x) { return x && x->world > 1 && x->rank >= 0; }
int xch_world(const PeerExchange* x) { return x ? x->world : 0; }
int xch_rank(const PeerExchange* x) { return x ? x->rank : -1; }

// publish the local lists of this step: own window + one flag store per peer
Status xch_publish(PeerExchange* x, const uint64_t* local_keys, const int32_t* ext_ids, uint32_t pos_base, size_t nq, size_t k,
                   cudaStream_t stream) {
  if (!xch_connected(x)) return Status::Err(9, "shard exchange is not connected");
  if (nq > x->cap_q || k > x->cap_k || nq * k > x->cap_q * x->cap_k)
    return Status::Err(6, "batch exceeds the exported exchange window (" + std::to_string(x->cap_q) + " queries x " +
                              std::to_string(x->cap_k) + ")");
  if ((size_t)x->world * k > (size_t)MERGE_ITEMS) return Status::Err(6, "world * k too large for the merge");
  const uint32_t step = ++x->step;
  const int parity = (int)(step & 1u);
  PeerPtrs pp;
  for (int p = 0; p < kMaxShardWorld; ++p) pp.win[p] = p < x->world ? x->peer[p] : nullptr;
  const size_t items = nq * k;
  const int pblocks = (int)std::min<size_t>((items + 255) / 256, 296);
  publish_kernel<<<pblocks, 256, 0, stream>>>(local_keys, ext_ids, pos_base, items, x->win, x->key_off(parity),
                                              x->id_off(parity), pp, x->rank, x->world, step);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return Status::Err(9, std::string("publish: ") + cudaGetErrorString(e));
  return Status::OK();
}

// merge queries [q_begin, q_begin + q_count) of all ranks' lists of the step just published
Status xch_merge(PeerExchange* x, size_t k, int finalize, size_t q_begin, size_t q_count, uint64_t* out_keys,
                 int32_t* out_ids, float* out_dists, int32_t* out_counts, cudaStream_t stream) {
  if (!xch_connected(x)) return Status::Err(9, "shard exchange is not connected");
  if (q_count == 0) return Status::OK();
  const uint32_t step = x->step;
  const int parity = (int)(step & 1u);
  PeerPtrs pp;
  for (int p = 0; p < kMaxShardWorld; ++p) pp.win[p] = p < x->world ? x->peer[p] : nullptr;
  int p2 = 1;
  while (p2 < x->world * (int)k) p2 <<= 1;
  int threads = std::max(32, std::min(256, p2 / 2));
  const size_t smem = (size_t)p2 * 12 + 16;
  cudaError_t e = cudaFuncSetAttribute(merge_peers_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MERGE_ITEMS * 12 + 16);
  if (e != cudaSuccess) return Status::Err(9, std::string("merge attr: ") + cudaGetErrorString(e));
  merge_peers_kernel<<<(unsigned)q_count, threads, smem, stream>>>(pp, x->win, x->key_off(parity), x->id_off(parity),
                                                                   x->world, step, (int)q_begin, (int)k, p2, finalize,
                                                                   out_keys, out_ids, out_dists, out_counts);
  e = cudaGetLastError();
  if (e != cudaSuccess) return Status::Err(9, std::string("merge: ") + cudaGetErrorString(e));
  return Status::OK();
}

// 1 if a wait in a merge kernel gave up (a peer never published); clears the word
bool xch_take_error(PeerExchange* x) {
  if (!x || !x->win) return false;
  unsigned v = 0;
  cudaSetDevice(x->device);
  if (cudaMemcpy(&v, x->win + 260, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  if (v) cudaMemset(x->win + 260, 0, 4);
  return v != 0;
}

}  // namespace nb200
