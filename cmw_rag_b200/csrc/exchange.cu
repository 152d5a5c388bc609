// exchange.cu -- K5x: the cross-shard candidate exchange fused with the merge, over NVLink peer memory.
//
// Row-sharded corpus: every rank holds a local top-k per query as (fp64 score, global
// id).  Instead of two NCCL all-gathers followed by the merge kernel, each rank's SEND kernel stores its
// candidates straight into the gather buffer of EVERY peer (st.global on cudaIpc-mapped peer pointers:
// NVLink 5 / NVSwitch writes), fences, and the last block publishes an epoch flag on each peer; the
// MERGE kernel of each rank spins on its own flags (one per source rank) and then reduces G*k -> k per
// query with the usual (score desc, id asc) rule.  The send kernel never waits, so every rank's flags
// are eventually published whatever the launch skew; buffers are double-buffered by epoch parity
// (a rank can only reach epoch n+2 after all peers have published n+1, i.e. finished reading n).
// The reference has no counterpart (single Chroma server).
//
// The merge kernel's wait is BOUNDED (a dead, late or out-of-step peer must not hang this GPU): it polls with
// __nanosleep back-off against %globaltimer and gives up after `timeout_ns`, marking every query
// CMW_FLAG_PEER_TIMEOUT; each sender also publishes the (B, k) it sent, and a mismatch is reported the same way.
//
// Peer buffer layout (one per rank, allocated by cmw_peer_alloc, zero-initialised):
//   [0, 2048)            header: uint32 flags[2][64] (parity, source rank), uint32 done[2], then at uint32
//                        offset 256: uint32 shape[2][64] = (B << 12 | k) as published by each source rank
//   [2048, ...)          2 parities x { f64 scores [G, Bmax, kmax] ; i64 ids [G, Bmax, kmax] ; i32 flags [G, Bmax] }
#include <string.h>

#include "common.cuh"

namespace cmw {

constexpr int kXMaxRanks = 64;
constexpr size_t kXHeaderBytes = 2048;
static inline size_t x_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct XPeers {
    uint8_t* buf[kXMaxRanks];
};

__device__ __forceinline__ uint32_t* x_flags(uint8_t* base, int parity) {
    return reinterpret_cast<uint32_t*>(base) + parity * kXMaxRanks;
}
__device__ __forceinline__ uint32_t* x_done(uint8_t* base, int parity) {
    return reinterpret_cast<uint32_t*>(base) + 2 * kXMaxRanks + parity;
}
__device__ __forceinline__ uint32_t* x_shape(uint8_t* base, int parity) {
    return reinterpret_cast<uint32_t*>(base) + 256 + parity * kXMaxRanks;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// publish (end of a send kernel, all threads): every block fences its peer stores system-wide; the last one to
// finish writes the shape word and then raises this rank's epoch flag on every peer
__device__ __forceinline__ void x_publish(const XPeers& peers, int G, int rank, int parity, uint32_t epoch,
                                          uint32_t shape) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* done = x_done(peers.buf[rank], parity);
        const uint32_t prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {
            *done = 0;  // ready for the next use of this parity
            __threadfence_system();
            for (int g = 0; g < G; ++g) *(volatile uint32_t*)(x_shape(peers.buf[g], parity) + rank) = shape;
            __threadfence_system();
            for (int g = 0; g < G; ++g) {
                volatile uint32_t* f = x_flags(peers.buf[g], parity) + rank;
                *f = epoch;
            }
            __threadfence_system();
        }
    }
}

// bounded wait (threads 0 .. G-1 of a block; the others pass through) until source rank threadIdx.x has published
// `epoch` with the expected shape word.  Returns false for a thread whose rank timed out or is out of step.
__device__ __forceinline__ bool x_wait(uint8_t* self, int G, int parity, uint32_t epoch, uint32_t shape,
                                       uint64_t timeout_ns) {
    if ((int)threadIdx.x >= G) return true;
    volatile uint32_t* f = x_flags(self, parity) + threadIdx.x;
    const uint64_t t0 = global_timer_ns();
    unsigned ns = 32;
    while (*f != epoch) {
        if (global_timer_ns() - t0 > timeout_ns) return false;
        __nanosleep(ns);
        if (ns < 2048) ns <<= 1;
    }
    __threadfence_system();
    return *(volatile uint32_t*)(x_shape(self, parity) + threadIdx.x) == shape;
}

__global__ void __launch_bounds__(256)
exchange_send_kernel(XPeers peers, int G, int rank, int B, int k, size_t slot_elems, size_t parity_bytes, int parity,
                     uint32_t epoch, const double* __restrict__ scores, const int64_t* __restrict__ ids,
                     const int32_t* __restrict__ flags, int max_batch) {
    const size_t n = (size_t)B * k;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int g = 0; g < G; ++g) {
        uint8_t* base = peers.buf[g] + kXHeaderBytes + (size_t)parity * parity_bytes;
        double* ds = reinterpret_cast<double*>(base) + (size_t)rank * slot_elems;
        int64_t* di = reinterpret_cast<int64_t*>(base + (size_t)G * slot_elems * sizeof(double)) + (size_t)rank * slot_elems;
        int32_t* df = reinterpret_cast<int32_t*>(base + (size_t)G * slot_elems * 16) + (size_t)rank * max_batch;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            ds[i] = scores[i];
            di[i] = ids[i];
            if (i < (size_t)B) df[i] = flags != nullptr ? flags[i] : 0;
        }
    }
    x_publish(peers, G, rank, parity, epoch, ((uint32_t)B << 12) | (uint32_t)k);
}

__global__ void __launch_bounds__(256)
exchange_merge_kernel(uint8_t* self, int G, int B, int k, int k_out, size_t slot_elems, size_t parity_bytes,
                      int parity, uint32_t epoch, uint64_t timeout_ns, float* __restrict__ out_scores,
                      int64_t* __restrict__ out_ids, double* __restrict__ out_scores64,
                      int32_t* __restrict__ out_flags, int max_batch) {
    extern __shared__ __align__(16) uint8_t x_smem[];
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    // wait -- for a bounded time -- until every source rank has published this epoch
    // every rank must exchange the same (B, k)
    if (!x_wait(self, G, parity, epoch, ((uint32_t)B << 12) | (uint32_t)k, timeout_ns)) atomicExch(&bad, 1);
    __syncthreads();
    __threadfence_system();
    const int b = blockIdx.x;
    if (bad) {
        for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
            out_scores[(size_t)b * k_out + j] = -INFINITY;
            out_ids[(size_t)b * k_out + j] = -1;
            if (out_scores64 != nullptr) out_scores64[(size_t)b * k_out + j] = -INFINITY;
        }
        if (threadIdx.x == 0 && out_flags != nullptr) out_flags[b] = CMW_FLAG_PEER_TIMEOUT;
        return;
    }
    if (threadIdx.x == 0 && out_flags != nullptr) {
        // OR of the shards' own flags (they travelled with the candidates)
        const int32_t* gf = reinterpret_cast<const int32_t*>(self + kXHeaderBytes + (size_t)parity * parity_bytes +
                                                             (size_t)G * slot_elems * 16);
        int f = 0;
        for (int g = 0; g < G; ++g) f |= *reinterpret_cast<const volatile int32_t*>(gf + (size_t)g * max_batch + b);
        out_flags[b] = f;
    }
    const int n = G * k;
    const int m = next_pow2(n < 2 ? 2 : n);
    uint64_t* hi = reinterpret_cast<uint64_t*>(x_smem);
    uint64_t* lo = hi + m;
    const uint8_t* base = self + kXHeaderBytes + (size_t)parity * parity_bytes;
    const double* gs = reinterpret_cast<const double*>(base);
    const int64_t* gi = reinterpret_cast<const int64_t*>(base + (size_t)G * slot_elems * sizeof(double));
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        uint64_t h = ~0ull, l = ~0ull;
        if (i < n) {
            const int g = i / k, j = i - g * k;
            const size_t src = (size_t)g * slot_elems + (size_t)b * k + j;
            // peers wrote these through NVLink: bypass L1 (volatile) -- they are fresh in L2 / HBM
            const int64_t id = *reinterpret_cast<const volatile int64_t*>(gi + src);
            const double s = *reinterpret_cast<const volatile double*>(gs + src);
            if (id >= 0 && s == s) {
                h = ~f64_orderable(s);
                l = (uint64_t)id;
            }
        }
        hi[i] = h;
        lo[i] = l;
    }
    bitonic_sort_u128(hi, lo, m);
    for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
        double s = -INFINITY;
        int64_t id = -1;
        if (j < n && !(hi[j] == ~0ull && lo[j] == ~0ull)) {
            s = f64_from_orderable(~hi[j]);
            id = (int64_t)lo[j];
        }
        out_scores[(size_t)b * k_out + j] = (float)s;
        out_ids[(size_t)b * k_out + j] = id;
        if (out_scores64 != nullptr) out_scores64[(size_t)b * k_out + j] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Generic peer all-gather: the two exchanges of the TWO-PHASE row-sharded search (sharded.py) over NVLink peer
// memory instead of NCCL.  Same buffer header and epoch-parity protocol as above.  A parity region holds G slots
// of `nbytes`, rank-major and contiguous -- exactly what an all-gather into one tensor leaves -- so the consumers
// (shard_kth_kernel, shard_merge_kernel) read the gathered data IN PLACE in this rank's peer buffer.  The send
// kernel stores this rank's slot into every peer's region and publishes; the wait kernel (one block) holds the
// stream until every rank's slot of this epoch has landed, for a bounded time: a dead or out-of-step peer sets
// CMW_FLAG_PEER_TIMEOUT in `status` (sticky; cmw_shard_merge_ex then flags every query) instead of hanging the GPU.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
peer_gather_send_kernel(XPeers peers, int G, int rank, size_t nbytes, size_t region_bytes, int parity, uint32_t epoch,
                        const uint8_t* __restrict__ src) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t off = kXHeaderBytes + (size_t)parity * region_bytes + (size_t)rank * nbytes;
    if ((nbytes & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        const size_t n = nbytes >> 4;
        for (size_t i = first; i < n; i += stride) {
            const uint4 v = __ldg(s4 + i);
            for (int g = 0; g < G; ++g) reinterpret_cast<uint4*>(peers.buf[g] + off)[i] = v;
        }
    } else {
        const uint32_t* s1 = reinterpret_cast<const uint32_t*>(src);
        const size_t n = nbytes >> 2;
        for (size_t i = first; i < n; i += stride) {
            const uint32_t v = __ldg(s1 + i);
            for (int g = 0; g < G; ++g) reinterpret_cast<uint32_t*>(peers.buf[g] + off)[i] = v;
        }
    }
    x_publish(peers, G, rank, parity, epoch, (uint32_t)(nbytes >> 2));
}

__global__ void __launch_bounds__(64)
peer_gather_wait_kernel(uint8_t* self, int G, int parity, uint32_t epoch, uint32_t shape, uint64_t timeout_ns,
                        int32_t* __restrict__ status) {
    if (!x_wait(self, G, parity, epoch, shape, timeout_ns)) atomicOr(status, CMW_FLAG_PEER_TIMEOUT);
    __threadfence_system();
}

}  // namespace cmw

using namespace cmw;

extern "C" {

size_t cmw_peer_buffer_bytes(int G, int max_batch, int max_k) {
    if (G < 1 || G > kXMaxRanks || max_batch < 1 || max_k < 1) return 0;
    return kXHeaderBytes + 2 * ((size_t)G * max_batch * max_k * 16 + (size_t)G * max_batch * 4 + 256);
}

int cmw_peer_alloc(int device, size_t bytes, void** dev_ptr, void* ipc_handle_out) {
    CMW_REQUIRE(dev_ptr != nullptr && ipc_handle_out != nullptr && bytes >= kXHeaderBytes, "cmw_peer_alloc: bad arguments");
    CMW_CUDA_OK(cudaSetDevice(device));
    void* p = nullptr;
    CMW_CUDA_OK(cudaMalloc(&p, bytes));
    CMW_CUDA_OK(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        set_error("cmw_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return -2;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    memcpy(ipc_handle_out, &h, sizeof(h));
    CMW_CUDA_OK(cudaDeviceSynchronize());
    *dev_ptr = p;
    return 0;
}

int cmw_peer_open(int device, const void* ipc_handle, void** dev_ptr) {
    CMW_REQUIRE(dev_ptr != nullptr && ipc_handle != nullptr, "cmw_peer_open: bad arguments");
    CMW_CUDA_OK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    CMW_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return 0;
}

int cmw_peer_close(void* dev_ptr) {
    if (dev_ptr) CMW_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

int cmw_peer_free(void* dev_ptr) {
    if (dev_ptr) CMW_CUDA_OK(cudaFree(dev_ptr));
    return 0;
}

size_t cmw_peer_gather_bytes(int G, size_t max_bytes_per_rank) {
    if (G < 1 || G > kXMaxRanks || max_bytes_per_rank < 4) return 0;
    return kXHeaderBytes + 2 * (size_t)G * x_align_up(max_bytes_per_rank, 256);
}

int cmw_peer_gather(void* const* peer_bufs_host, int G, int rank, size_t max_bytes_per_rank, const void* src_dev,
                    size_t nbytes, uint32_t epoch, int timeout_ms, int32_t* status_dev, void** gathered_dev_out,
                    void* stream_v) {
    CMW_REQUIRE(peer_bufs_host && src_dev && status_dev && gathered_dev_out, "cmw_peer_gather: NULL argument");
    CMW_REQUIRE(G >= 1 && G <= kXMaxRanks && rank >= 0 && rank < G, "cmw_peer_gather: bad rank/world");
    CMW_REQUIRE(nbytes >= 4 && (nbytes & 3) == 0 && nbytes <= max_bytes_per_rank && (nbytes >> 2) < 0xffffffffull,
                "cmw_peer_gather: nbytes = %zu must be a multiple of 4 and at most %zu", nbytes, max_bytes_per_rank);
    CMW_REQUIRE((reinterpret_cast<uintptr_t>(src_dev) & 15) == 0, "cmw_peer_gather: src_dev must be 16-byte aligned");
    CMW_REQUIRE(epoch != 0, "cmw_peer_gather: epoch must be non-zero");
    cudaStream_t stream = (cudaStream_t)stream_v;
    XPeers peers;
    for (int g = 0; g < G; ++g) {
        CMW_REQUIRE(peer_bufs_host[g] != nullptr, "cmw_peer_gather: peer buffer %d is NULL", g);
        peers.buf[g] = reinterpret_cast<uint8_t*>(peer_bufs_host[g]);
    }
    const size_t region_bytes = (size_t)G * x_align_up(max_bytes_per_rank, 256);
    const int parity = (int)(epoch & 1u);
    const uint64_t timeout_ns = (uint64_t)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
    const size_t vecs = (nbytes & 15) == 0 ? nbytes >> 4 : nbytes >> 2;
    int blocks = (int)((vecs + 255) / 256);
    if (blocks > 296) blocks = 296;
    if (blocks < 1) blocks = 1;
    peer_gather_send_kernel<<<blocks, 256, 0, stream>>>(peers, G, rank, nbytes, region_bytes, parity, epoch,
                                                        reinterpret_cast<const uint8_t*>(src_dev));
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    peer_gather_wait_kernel<<<1, 64, 0, stream>>>(peers.buf[rank], G, parity, epoch, (uint32_t)(nbytes >> 2),
                                                  timeout_ns, status_dev);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    *gathered_dev_out = peers.buf[rank] + kXHeaderBytes + (size_t)parity * region_bytes;
    return 0;
}

int cmw_exchange_merge(void* const* peer_bufs_host, int G, int rank, int max_batch, int max_k, int B, int k,
                       int k_out, uint32_t epoch, const double* scores64_local_dev, const int64_t* ids_local_dev,
                       float* out_scores_dev, int64_t* out_ids_dev, double* out_scores64_dev, void* stream_v) {
    return cmw_exchange_merge_ex(peer_bufs_host, G, rank, max_batch, max_k, B, k, k_out, epoch, scores64_local_dev,
                                 ids_local_dev, nullptr, out_scores_dev, out_ids_dev, out_scores64_dev, nullptr, 0,
                                 stream_v);
}

int cmw_exchange_merge_ex(void* const* peer_bufs_host, int G, int rank, int max_batch, int max_k, int B, int k,
                          int k_out, uint32_t epoch, const double* scores64_local_dev, const int64_t* ids_local_dev,
                          const int32_t* flags_local_dev, float* out_scores_dev, int64_t* out_ids_dev,
                          double* out_scores64_dev, int32_t* out_flags_dev, int timeout_ms, void* stream_v) {
    CMW_REQUIRE(peer_bufs_host && scores64_local_dev && ids_local_dev && out_scores_dev && out_ids_dev,
                "cmw_exchange_merge: NULL argument");
    CMW_REQUIRE(G >= 1 && G <= kXMaxRanks && rank >= 0 && rank < G, "cmw_exchange_merge: bad rank/world");
    CMW_REQUIRE(B >= 1 && B <= max_batch && k >= 1 && k <= max_k && k_out >= 1, "cmw_exchange_merge: bad sizes");
    CMW_REQUIRE(B < (1 << 20) && k < (1 << 12), "cmw_exchange_merge: B or k too large for the shape word");
    const uint64_t timeout_ns = (uint64_t)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
    CMW_REQUIRE(epoch != 0, "cmw_exchange_merge: epoch must be non-zero");
    const int n = G * k;
    CMW_REQUIRE(n <= 8192, "cmw_exchange_merge: G*k = %d exceeds 8192", n);
    cudaStream_t stream = (cudaStream_t)stream_v;
    XPeers peers;
    for (int g = 0; g < G; ++g) {
        CMW_REQUIRE(peer_bufs_host[g] != nullptr, "cmw_exchange_merge: peer buffer %d is NULL", g);
        peers.buf[g] = reinterpret_cast<uint8_t*>(peer_bufs_host[g]);
    }
    const size_t slot_elems = (size_t)max_batch * max_k;
    const size_t parity_bytes = (size_t)G * slot_elems * 16 + (size_t)G * max_batch * 4 + 256;
    const int parity = (int)(epoch & 1u);
    const size_t total = (size_t)B * k;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 296) blocks = 296;
    exchange_send_kernel<<<blocks, 256, 0, stream>>>(peers, G, rank, B, k, slot_elems, parity_bytes, parity, epoch,
                                                     scores64_local_dev, ids_local_dev, flags_local_dev, max_batch);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    const size_t smem = (size_t)next_pow2_host(n < 2 ? 2 : n) * 16;
    static SmemAttrCache smem_set;
    if (smem > 48 * 1024 && smem_set.needs(smem)) {
        CMW_CUDA_OK(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set.done(smem);
    }
    exchange_merge_kernel<<<B, 256, smem, stream>>>(peers.buf[rank], G, B, k, k_out, slot_elems, parity_bytes, parity,
                                                    epoch, timeout_ns, out_scores_dev, out_ids_dev, out_scores64_dev,
                                                    out_flags_dev, max_batch);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
