// Peer-memory exchange primitives for one-process-per-GPU BESS (one entity
// shard per GPU over NVLink 5 / NVSwitch).
//
// The reference exchanges the n x n triple blocks with a balanced AllToAll
// (bess.py:348-350; autograd gives the reverse AllToAll of the gradients, and
// PopTorch all-reduces the replicated relation gradient).  Here the gather
// kernel (rows.cu: bess_gather_route) already stores every tail / negative row
// straight into the destination GPU's receive buffer through a peer-mapped
// pointer, so the forward "collective" is just that kernel plus a flag
// handshake; the gradient blocks and the relation partials are pushed with
// vectorised remote stores.  No NCCL call is left inside the step, which makes
// the whole step capturable in a CUDA graph.
//
// Handshake: every rank owns, per channel, a row of n int32 flags (symmetric
// memory) and a local int32 sequence counter.  signal: seq = ++counter, then
// flag[my_rank] on every peer := seq with release semantics at system scope.
// wait: spin (acquire, system scope) until all n local flags >= counter.
// Waits are bounded in wall-clock time (caller-chosen, minutes by default): a peer that never
// arrives traps the launch instead of hanging the GPU for ever.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace bess {

struct PeerPtrs {
  void* p[BESS_MAX_SHARD];
};

BESS_D void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
BESS_D int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void peer_signal_kernel(int32_t* counter, PeerPtrs peer_flags, int my_rank, int n) {
  __shared__ int32_t seq;
  if (threadIdx.x == 0) {
    seq = *counter + 1;
    *counter = seq;
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {
    __threadfence_system();  // everything this GPU stored before (previous kernels) is visible first
    st_release_sys(reinterpret_cast<int32_t*>(peer_flags.p[threadIdx.x]) + my_rank, seq);
  }
}

BESS_D uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Bounded spin on wall-clock time (%globaltimer, nanoseconds: independent of the SM clock).
// timeout_ns <= 0 waits for ever.  On expiry the launch reports which peer is missing and
// traps: continuing would score / update with rows that never arrived.
__global__ void peer_wait_kernel(const int32_t* counter, const int32_t* my_flags, int n,
                                 int64_t timeout_ns) {
  const int32_t expected = *counter;
  if ((int)threadIdx.x < n) {
    const uint64_t t0 = globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(my_flags + threadIdx.x) < expected) {
      if (timeout_ns > 0 && (++spins & 1023u) == 0 &&
          globaltimer_ns() - t0 > (uint64_t)timeout_ns) {
        printf("besskge_b200 peer_wait: flag of rank %d stuck at %d, expected %d after %lld ms "
               "(BESS_PEER_TIMEOUT_S)\n", (int)threadIdx.x, ld_acquire_sys(my_flags + threadIdx.x),
               expected, (long long)(timeout_ns / 1000000));
        __trap();
      }
    }
  }
  __threadfence_system();
}

// Stage stamp: one thread stores %globaltimer (ns).  Placed between the kernels of a captured step
// it gives per-stage device timings of a multi-rank run without a profiler (bench --stage-timing).
__global__ void stamp_kernel(unsigned long long* out) { *out = globaltimer_ns(); }

// block (j, *) copies src + j * src_stride_bytes -> dst[j], bytes_each bytes (16-byte multiples).
// Four 16-byte loads are in flight per thread before the first remote store: the launcher keeps
// the grid at ~2 blocks per SM so that a persistent GEMM on another stream still finds room on
// every SM (a grid that fills all 2048 thread slots serialises the two), and the unrolling is
// what keeps NVLink full from that small footprint.
__global__ void __launch_bounds__(256) peer_push_kernel(const uint8_t* __restrict__ src,
                                                        int64_t src_stride_bytes, PeerPtrs dst,
                                                        int64_t bytes_each) {
  const int j = blockIdx.y;
  const uint4* s = reinterpret_cast<const uint4*>(src + (int64_t)j * src_stride_bytes);
  uint4* d = reinterpret_cast<uint4*>(dst.p[j]);
  const int64_t n16 = bytes_each >> 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = ld_stream(s + i), b = ld_stream(s + i + stride);
    const uint4 c = ld_stream(s + i + 2 * stride), e = ld_stream(s + i + 3 * stride);
    st_stream(d + i, a);
    st_stream(d + i + stride, b);
    st_stream(d + i + 2 * stride, c);
    st_stream(d + i + 3 * stride, e);
  }
  for (; i < n16; i += stride) st_stream(d + i, ld_stream(s + i));
}

// out[i] = scale * sum_j slots[j * count + i], j ascending (same order on every rank)
__global__ void peer_reduce_kernel(const float* __restrict__ slots, int n, int64_t count, float scale,
                                   float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc += slots[(int64_t)j * count + i];
    out[i] = acc * scale;
  }
}

}  // namespace bess

using namespace bess;

static int fill_ptrs(PeerPtrs& pp, void* const* ptrs, int n, const char* what) {
  if (n < 1 || n > BESS_MAX_SHARD) {
    bess_set_error("%s: n=%d out of range (1..%d)", what, n, BESS_MAX_SHARD);
    return BESS_ERR_INVALID_ARG;
  }
  for (int i = 0; i < BESS_MAX_SHARD; ++i) pp.p[i] = i < n ? ptrs[i] : nullptr;
  return BESS_OK;
}

extern "C" int bess_peer_signal(int32_t* counter, void* const* peer_flags, int my_rank, int n,
                                void* stream) {
  PeerPtrs pp;
  if (int e = fill_ptrs(pp, peer_flags, n, "bess_peer_signal")) return e;
  peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter, pp, my_rank, n);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_peer_wait(const int32_t* counter, const int32_t* my_flags, int n,
                              int64_t timeout_ms, void* stream) {
  BESS_CHECK_ARG(n >= 1 && n <= 32, "bess_peer_wait: n=%d out of range", n);
  peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counter, my_flags, n,
                                                      timeout_ms * 1000000LL);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_stamp(uint64_t* out, void* stream) {
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(out));
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

// blocks per SM of the push grid (all destinations together); BESS_PUSH_BLOCKS_PER_SM overrides
static int push_blocks_per_sm() {
  static int v = [] {
    const char* e = getenv("BESS_PUSH_BLOCKS_PER_SM");
    const int x = e != nullptr ? atoi(e) : 2;
    return x < 1 ? 1 : (x > 8 ? 8 : x);
  }();
  return v;
}

extern "C" int bess_peer_push(const void* src, int64_t src_stride_bytes, void* const* dst, int n,
                              int64_t bytes_each, void* stream) {
  if (bytes_each <= 0) return BESS_OK;
  PeerPtrs pp;
  if (int e = fill_ptrs(pp, dst, n, "bess_peer_push")) return e;
  BESS_CHECK_ARG(bytes_each % 16 == 0 && src_stride_bytes % 16 == 0 && ((uintptr_t)src & 15) == 0,
                 "bess_peer_push: 16-byte alignment required");
  const int64_t n16 = bytes_each >> 4;
  int bx = (int)((n16 + 255) / 256);
  const int cap = (push_blocks_per_sm() * kNumSM + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  peer_push_kernel<<<dim3(bx, n), 256, 0, (cudaStream_t)stream>>>((const uint8_t*)src, src_stride_bytes,
                                                                 pp, bytes_each);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

extern "C" int bess_peer_copy(const void* src, void* dst, int64_t bytes, void* stream) {
  if (bytes <= 0) return BESS_OK;
  BESS_CHECK_ARG(src != nullptr && dst != nullptr, "bess_peer_copy: null pointer");
  const cudaError_t e =
      cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    bess_set_error("bess_peer_copy: %s", cudaGetErrorString(e));
    return BESS_ERR_CUDA;
  }
  return BESS_OK;  // a memcpy node, not a kernel: bess_launch_count() does not move
}

extern "C" int bess_peer_reduce(const float* slots, int n, int64_t count, float scale, float* out,
                                void* stream) {
  if (count <= 0) return BESS_OK;
  int blocks = (int)((count + 255) / 256);
  if (blocks > 4 * kNumSM) blocks = 4 * kNumSM;
  peer_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(slots, n, count, scale, out);
  BESS_CHECK_LAUNCH();
  return BESS_OK;
}

// ---------------------------------------------------------------------------
// Peer-mapped buffers without any framework: cudaMalloc + CUDA IPC handles.
// Every rank allocates its receive buffer with bess_peer_alloc, exports a 64-byte handle,
// the handles travel through whatever side channel the host has (torch.distributed
// all_gather_object here) and bess_peer_import maps each peer's buffer into this process.
// Works between GPUs of one node (peer access over NVLink is enabled lazily by the driver)
// and between two processes that share ONE GPU — which is how the 2-rank protocol test runs
// on a single-GPU box.  torch's symmetric-memory allocator (the default transport on
// multi-GPU nodes) refuses the latter.
// ---------------------------------------------------------------------------
#define BESS_CUDA_TRY(expr, what)                                                   \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      bess_set_error("%s: %s", what, cudaGetErrorString(_e));                       \
      return BESS_ERR_CUDA;                                                         \
    }                                                                               \
  } while (0)

extern "C" int bess_peer_alloc(int64_t bytes, void** ptr) {
  BESS_CHECK_ARG(bytes > 0 && ptr != nullptr, "bess_peer_alloc: bytes > 0 and an out pointer required");
  BESS_CUDA_TRY(cudaMalloc(ptr, (size_t)bytes), "bess_peer_alloc (cudaMalloc)");
  BESS_CUDA_TRY(cudaMemset(*ptr, 0, (size_t)bytes), "bess_peer_alloc (cudaMemset)");
  BESS_CUDA_TRY(cudaDeviceSynchronize(), "bess_peer_alloc (sync)");
  return BESS_OK;
}

extern "C" int bess_peer_free(void* ptr) {
  if (ptr != nullptr) BESS_CUDA_TRY(cudaFree(ptr), "bess_peer_free");
  return BESS_OK;
}

extern "C" int bess_peer_export(void* ptr, void* handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == BESS_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  BESS_CUDA_TRY(cudaIpcGetMemHandle(&h, ptr), "bess_peer_export (cudaIpcGetMemHandle)");
  memcpy(handle_out, &h, sizeof(h));
  return BESS_OK;
}

extern "C" int bess_peer_import(const void* handle, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  BESS_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess),
                "bess_peer_import (cudaIpcOpenMemHandle)");
  return BESS_OK;
}

extern "C" int bess_peer_unmap(void* ptr) {
  if (ptr != nullptr) BESS_CUDA_TRY(cudaIpcCloseMemHandle(ptr), "bess_peer_unmap");
  return BESS_OK;
}
