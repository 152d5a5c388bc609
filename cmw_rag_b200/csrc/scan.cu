// scan.cu -- K1: streaming dot-product filter, 1-4 queries per pass over the corpus (HBM-bound regime).
//
// Stands in for the HNSW walk behind collection.query() (rag_engine/storage/vector_store.py:59-63
// of the reference) for single queries -- but exact: every live row is scored.
//
// Design (B200): one persistent CTA per SM.  Warp 0 is the producer: one elected lane streams the
// slab with 1-D bulk async copies (TMA engine, cp.async.bulk + mbarrier complete_tx, L2 evict-first)
// into a 4-deep ring of 48 KB stages, so ~192 KB per SM are in flight independently of register
// pressure or occupancy.  Eight consumer warps each take one row of a stage at a time: conflict-free
// 128-bit LDS, fp32 FMA against the query held in registers, xor-shuffle reduction, multiply by the
// per-row multiplier (1/|c|, |c| or 1; NaN for tombstoned rows) that travels in the same stage.
// A row whose score passes the pool threshold is appended to the query's candidate pool with one
// global atomic; in the first ("dense") slab every row is written to its own slot instead.
// The score matrix is never materialised.
//
// Algorithmic bytes per row: D * sizeof(elt) + 4 (multiplier).
#include "common.cuh"
#include "ptx.cuh"

namespace cmw {

constexpr int kScanConsumerWarps = 8;
constexpr int kScanThreads = (kScanConsumerWarps + 1) * 32;
constexpr int kScanStageTarget = 48 * 1024;
constexpr int kScanMaxStages = 4;

struct ScanParams {
    const uint8_t* rows;
    const float* row_mul;
    const float* q;
    int64_t row_begin, row_end;
    int dim;
    int row_bytes;
    int rows_per_stage;
    int stage_bytes;  // rows region + multiplier region, 128-byte aligned
    int mul_offset;   // byte offset of the multiplier region inside a stage
    int nstages;
    int dense;
    float* dense_scores;
    int32_t* dense_ids;
    int dense_stride;
    float* pool_scores;
    int32_t* pool_ids;
    int32_t* pool_cnt;
    const float* pool_thr;
};

template <typename ELT>
struct EltTraits;
template <>
struct EltTraits<float> {
    static constexpr int kPerVec = 4;  // elements per 128-bit load
};
template <>
struct EltTraits<__nv_bfloat16> {
    static constexpr int kPerVec = 8;
};
template <>
struct EltTraits<__half> {
    static constexpr int kPerVec = 8;
};

template <int NQ>
__device__ __forceinline__ void fma_vec(const float4& v, const float4 (&qv)[NQ], float (&acc)[NQ]) {
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        acc[i] = fmaf(v.x, qv[i].x, acc[i]);
        acc[i] = fmaf(v.y, qv[i].y, acc[i]);
        acc[i] = fmaf(v.z, qv[i].z, acc[i]);
        acc[i] = fmaf(v.w, qv[i].w, acc[i]);
    }
}

// eight 16-bit tile elements -> two float4 (bf16: a shift; fp16: cvt)
template <typename ELT>
__device__ __forceinline__ void unpack_x8(const uint4& u, float4& a, float4& b);
template <>
__device__ __forceinline__ void unpack_x8<float>(const uint4&, float4&, float4&) {}
template <>
__device__ __forceinline__ void unpack_x8<__half>(const uint4& u, float4& a, float4& b) {
    const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    const float2 p2 = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
    const float2 p3 = __half22float2(*reinterpret_cast<const __half2*>(&u.w));
    a = make_float4(p0.x, p0.y, p1.x, p1.y);
    b = make_float4(p2.x, p2.y, p3.x, p3.y);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float4& a, float4& b) {
    a.x = __uint_as_float(u.x << 16);
    a.y = __uint_as_float(u.x & 0xffff0000u);
    a.z = __uint_as_float(u.y << 16);
    a.w = __uint_as_float(u.y & 0xffff0000u);
    b.x = __uint_as_float(u.z << 16);
    b.y = __uint_as_float(u.z & 0xffff0000u);
    b.z = __uint_as_float(u.w << 16);
    b.w = __uint_as_float(u.w & 0xffff0000u);
}
template <>
__device__ __forceinline__ void unpack_x8<__nv_bfloat16>(const uint4& u, float4& a, float4& b) {
    unpack_bf16x8(u, a, b);
}

// CHUNKS > 0: dim == CHUNKS * 32 * kPerVec and the query lives in registers.
// CHUNKS == 0: any dim (multiple of 8); the query is read from shared memory.
template <typename ELT, int NQ, int CHUNKS>
__global__ void __launch_bounds__(kScanThreads, 1) scan_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int PV = EltTraits<ELT>::kPerVec;
    constexpr int F4 = PV / 4;  // float4 of query per corpus vector (1 for fp32, 2 for bf16)
    uint8_t* stage_base = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.nstages * p.stage_bytes);
    uint64_t* empty = full + kScanMaxStages;
    float* q_smem = reinterpret_cast<float*>(empty + kScanMaxStages);  // only when CHUNKS == 0

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int R = p.rows_per_stage;
    const int64_t nrows = p.row_end - p.row_begin;
    const int64_t total_stages = (nrows + R - 1) / R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.nstages; ++s) {
            ptx::mbar_init(&full[s], 1);
            ptx::mbar_init(&empty[s], kScanConsumerWarps);
        }
        ptx::fence_barrier_init();
    }
    if (CHUNKS == 0) {
        for (int i = threadIdx.x; i < NQ * p.dim; i += blockDim.x) q_smem[i] = p.q[i];
    }
    __syncthreads();

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            const uint64_t policy = ptx::l2_policy_evict_first();
            int it = 0;
            for (int64_t s = blockIdx.x; s < total_stages; s += gridDim.x, ++it) {
                const int slot = it % p.nstages;
                const uint32_t phase = (uint32_t)(it / p.nstages) & 1u;
                ptx::mbar_wait(&empty[slot], phase ^ 1u);
                const int64_t r0 = p.row_begin + s * R;
                const int64_t left = p.row_end - r0;
                const int nr = left < R ? (int)left : R;
                const uint32_t bytes_rows = (uint32_t)nr * (uint32_t)p.row_bytes;
                const uint32_t bytes_mul = (uint32_t)((nr + 3) & ~3) * 4u;
                uint8_t* dst = stage_base + (size_t)slot * p.stage_bytes;
                ptx::mbar_arrive_expect_tx(&full[slot], bytes_rows + bytes_mul);
                ptx::bulk_g2s(dst, p.rows + (size_t)r0 * p.row_bytes, bytes_rows, &full[slot], policy);
                ptx::bulk_g2s(dst + p.mul_offset, p.row_mul + r0, bytes_mul, &full[slot], policy);
            }
        }
        return;
    }

    // ===== consumers =====
    const int cw = warp - 1;
    float4 qreg[NQ][CHUNKS > 0 ? CHUNKS * F4 : 1];
    if (CHUNKS > 0) {
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const float4* qp = reinterpret_cast<const float4*>(p.q + (size_t)i * p.dim);
#pragma unroll
            for (int j = 0; j < CHUNKS; ++j) {
#pragma unroll
                for (int h = 0; h < F4; ++h) qreg[i][j * F4 + h] = __ldg(qp + (lane + 32 * j) * F4 + h);
            }
        }
    }
    float thr[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) thr[i] = p.dense ? 0.f : p.pool_thr[i];

    const int nvec = p.dim / PV;
    int it = 0;
    for (int64_t s = blockIdx.x; s < total_stages; s += gridDim.x, ++it) {
        const int slot = it % p.nstages;
        const uint32_t phase = (uint32_t)(it / p.nstages) & 1u;
        ptx::mbar_wait(&full[slot], phase);
        const int64_t r0 = p.row_begin + s * R;
        const int64_t left = p.row_end - r0;
        const int nr = left < R ? (int)left : R;
        const uint8_t* st = stage_base + (size_t)slot * p.stage_bytes;
        const float* mul = reinterpret_cast<const float*>(st + p.mul_offset);
        for (int rr = cw; rr < nr; rr += kScanConsumerWarps) {
            float acc[NQ];
#pragma unroll
            for (int i = 0; i < NQ; ++i) acc[i] = 0.f;
            const uint4* rowp = reinterpret_cast<const uint4*>(st + (size_t)rr * p.row_bytes);
            if (CHUNKS > 0) {
#pragma unroll
                for (int j = 0; j < CHUNKS; ++j) {
                    const uint4 u = rowp[lane + 32 * j];
                    if (F4 == 1) {
                        float4 v = make_float4(__uint_as_float(u.x), __uint_as_float(u.y),
                                               __uint_as_float(u.z), __uint_as_float(u.w));
                        float4 qv[NQ];
#pragma unroll
                        for (int i = 0; i < NQ; ++i) qv[i] = qreg[i][j];
                        fma_vec<NQ>(v, qv, acc);
                    } else {
                        float4 a, b;
                        unpack_x8<ELT>(u, a, b);
                        float4 qa[NQ], qb[NQ];
#pragma unroll
                        for (int i = 0; i < NQ; ++i) {
                            qa[i] = qreg[i][j * F4];
                            qb[i] = qreg[i][j * F4 + (F4 - 1)];
                        }
                        fma_vec<NQ>(a, qa, acc);
                        fma_vec<NQ>(b, qb, acc);
                    }
                }
            } else {
                for (int c = lane; c < nvec; c += 32) {
                    const uint4 u = rowp[c];
                    if (F4 == 1) {
                        float4 v = make_float4(__uint_as_float(u.x), __uint_as_float(u.y),
                                               __uint_as_float(u.z), __uint_as_float(u.w));
                        float4 qv[NQ];
#pragma unroll
                        for (int i = 0; i < NQ; ++i)
                            qv[i] = reinterpret_cast<const float4*>(q_smem + (size_t)i * p.dim)[c];
                        fma_vec<NQ>(v, qv, acc);
                    } else {
                        float4 a, b;
                        unpack_x8<ELT>(u, a, b);
                        float4 qa[NQ], qb[NQ];
#pragma unroll
                        for (int i = 0; i < NQ; ++i) {
                            const float4* qp = reinterpret_cast<const float4*>(q_smem + (size_t)i * p.dim);
                            qa[i] = qp[2 * c];
                            qb[i] = qp[2 * c + 1];
                        }
                        fma_vec<NQ>(a, qa, acc);
                        fma_vec<NQ>(b, qb, acc);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NQ; ++i) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
            }
            if (lane == 0) {
                const float m = mul[rr];
                const int64_t r = r0 + rr;
#pragma unroll
                for (int i = 0; i < NQ; ++i) {
                    const float sc = acc[i] * m;
                    if (p.dense) {
                        const size_t pos = (size_t)i * (size_t)p.dense_stride + (size_t)(r - p.row_begin);
                        p.dense_scores[pos] = (sc == sc) ? sc : -INFINITY;
                        p.dense_ids[pos] = (int32_t)r;
                    } else if (sc >= thr[i]) {
                        const int pos = atomicAdd(p.pool_cnt + i, 1);
                        if (pos < kPoolCap) {
                            p.pool_scores[(size_t)i * kPoolCap + pos] = sc;
                            p.pool_ids[(size_t)i * kPoolCap + pos] = (int32_t)r;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty[slot]);
    }
}

template <typename ELT, int NQ, int CHUNKS>
static int launch_one(const ScanParams& p, int grid, size_t smem, cudaStream_t stream) {
    auto kern = scan_kernel<ELT, NQ, CHUNKS>;
    static SmemAttrCache smem_set;  // per instantiation
    if (smem_set.needs(smem)) {
        CMW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set.done(smem);
    }
    kern<<<grid, kScanThreads, smem, stream>>>(p);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename ELT, int NQ>
static int dispatch_chunks(const ScanParams& p, int grid, size_t smem, cudaStream_t stream) {
    constexpr int PV = EltTraits<ELT>::kPerVec;
    // 3-4 queries: NQ * D / 32 query floats per lane no longer fit the register file next to the row vectors
    // (ptxas spills from NQ = 3 at D = 1536), so the queries are read from shared memory: ~30 KB of LDS per
    // 6 KB row, still just inside the SM's shared-memory bandwidth at the HBM feed rate
    if constexpr (NQ > 2) {
        return launch_one<ELT, NQ, 0>(p, grid, smem, stream);
    } else {
        if (p.dim == 1536) return launch_one<ELT, NQ, 1536 / (32 * PV)>(p, grid, smem, stream);
        if (p.dim == 1024) return launch_one<ELT, NQ, 1024 / (32 * PV)>(p, grid, smem, stream);
        if (p.dim == 768 && PV == 4) return launch_one<ELT, NQ, 768 / 128>(p, grid, smem, stream);
        return launch_one<ELT, NQ, 0>(p, grid, smem, stream);
    }
}

int launch_scan(const ScanArgs& a, cudaStream_t stream) {
    CMW_REQUIRE(a.nq >= 1 && a.nq <= kScanMaxQueries, "launch_scan: nq must be in [1, %d]", kScanMaxQueries);
    CMW_REQUIRE(a.row_begin % 4 == 0, "launch_scan: slab start must be a multiple of 4");
    if (a.row_end <= a.row_begin) return 0;
    ScanParams p;
    p.rows = reinterpret_cast<const uint8_t*>(a.rows);
    p.row_mul = a.row_mul;
    p.q = a.q;
    p.row_begin = a.row_begin;
    p.row_end = a.row_end;
    p.dim = a.dim;
    p.row_bytes = a.dim * a.elt_bytes;
    int R = (kScanStageTarget / p.row_bytes) & ~3;
    if (R < 4) R = 4;
    p.rows_per_stage = R;
    p.mul_offset = R * p.row_bytes;  // multiple of 16
    p.stage_bytes = (p.mul_offset + R * 4 + 127) & ~127;
    const size_t tail = 2 * kScanMaxStages * sizeof(uint64_t) + (size_t)a.nq * a.dim * sizeof(float) + 64;
    int nst = (int)((200 * 1024 - tail) / (size_t)p.stage_bytes);
    if (nst > kScanMaxStages) nst = kScanMaxStages;
    CMW_REQUIRE(nst >= 2, "launch_scan: dim %d too large for the staging ring", a.dim);
    p.nstages = nst;
    p.dense = a.dense;
    p.dense_scores = a.wide_scores ? a.wide_scores : a.pool.scores;
    p.dense_ids = a.wide_scores ? a.wide_ids : a.pool.ids;
    p.dense_stride = a.wide_scores ? a.wide_stride : kPoolCap;
    p.pool_scores = a.pool.scores;
    p.pool_ids = a.pool.ids;
    p.pool_cnt = a.pool.cnt;
    p.pool_thr = a.pool.thr;
    if (a.dense)
        CMW_REQUIRE(a.row_end - a.row_begin <= (a.wide_scores ? a.wide_stride : kPoolCap),
                    "launch_scan: dense slab larger than its destination");
    const size_t smem = (size_t)nst * p.stage_bytes + tail;
    const int64_t total_stages = (a.row_end - a.row_begin + R - 1) / R;
    int grid = a.sm_count;
    if ((int64_t)grid > total_stages) grid = (int)total_stages;
    if (a.elt_bytes == 4) {
        switch (a.nq) {
            case 1: return dispatch_chunks<float, 1>(p, grid, smem, stream);
            case 2: return dispatch_chunks<float, 2>(p, grid, smem, stream);
            case 3: return dispatch_chunks<float, 3>(p, grid, smem, stream);
            default: return dispatch_chunks<float, 4>(p, grid, smem, stream);
        }
    }
    if (a.half_tiles) {
        switch (a.nq) {
            case 1: return dispatch_chunks<__half, 1>(p, grid, smem, stream);
            case 2: return dispatch_chunks<__half, 2>(p, grid, smem, stream);
            case 3: return dispatch_chunks<__half, 3>(p, grid, smem, stream);
            default: return dispatch_chunks<__half, 4>(p, grid, smem, stream);
        }
    }
    switch (a.nq) {
        case 1: return dispatch_chunks<__nv_bfloat16, 1>(p, grid, smem, stream);
        case 2: return dispatch_chunks<__nv_bfloat16, 2>(p, grid, smem, stream);
        case 3: return dispatch_chunks<__nv_bfloat16, 3>(p, grid, smem, stream);
        default: return dispatch_chunks<__nv_bfloat16, 4>(p, grid, smem, stream);
    }
}

}  // namespace cmw
