// gemm_common.cuh -- pieces shared by the two K2 kernels (1-CTA in gemm.cu, CTA-pair in gemm2.cu):
// parameters, the shared-memory operand descriptor and the fused top-k admission epilogue.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace cmw {

constexpr int kTileM = 128;          // corpus rows per CTA per item (= TMEM lanes)
constexpr int kBlockK = 64;          // 16-bit elements per k-block = one 128-byte swizzle atom (fp32: 32, GemmParams::kb_elems)
constexpr int kUmmaK = 16;
constexpr int kMaxNT = 256;
constexpr int kABytes = kTileM * kBlockK * 2;  // 16 KB
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;      // TMEM columns per accumulator stage
constexpr int kStageCap = 512;       // staged survivors per epilogue warp (8 bytes each), 1-CTA kernel
constexpr int kStageCap2 = 256;      // same, CTA-pair kernel (8 epilogue warps, 128 columns each)
constexpr int kAggregateFrom = 64;   // staged survivors from which a flush reserves pool slots per run of one query

struct GemmParams {
    int dim;
    int num_kb;           // ceil(dim / 64)
    int nt;               // queries per group (multiple of 16, <= 256)
    int n_groups;
    int batch;            // real queries
    int64_t row_begin, row_end;
    int n_tiles;          // row tiles (1-CTA: 128 rows, CTA pair: 256 rows)
    int nstages;
    int stage_bytes;
    int dense;
    float* dense_scores;  // dense slab: where the scores go ([query, dense_stride]: the pools, or the wide scratch)
    int32_t* dense_ids;
    int dense_stride;
    // Scan order.  Tile g of the store (g = tile_begin + the launch's tile index; a tile = 128 rows, or 256 for a
    // CTA pair) covers the rows of tile (g * perm_mul) mod perm_tiles: a stride permutation with perm_mul
    // coprime to perm_tiles and close to perm_tiles / golden ratio, so that ANY range of consecutive tile
    // indices -- every slab -- is a low-discrepancy sample of the whole corpus and the admission thresholds
    // it yields are valid estimates whatever order the corpus is stored in.  perm_mul = 0: identity.
    int64_t tile_begin;
    int64_t perm_mul;
    int64_t perm_tiles;
    int dynamic;          // CTA-pair kernel: 1 = work items handed out by cluster launch control
    int tf32;             // 1-CTA kernel: operands are fp32 rows read by kind::tf32 MMAs (K = 8 per instruction)
    int kb_elems;         // elements per 128-byte k-block: 64 (16-bit tiles) or 32 (fp32)
    uint32_t idesc;
    const float* row_mul;
    float* pool_scores;
    int32_t* pool_ids;
    int32_t* pool_cnt;
    const float* pool_thr;
};

// width of one query group of a padded batch (api.cu: pad_batch): the whole batch up to 256 queries, else the
// batch split evenly over ceil(bpad / 256) groups
__host__ __device__ inline int gemm_group_width(int bpad) {
    if (bpad <= kMaxNT) return bpad;
    const int groups = (bpad + kMaxNT - 1) / kMaxNT;
    return bpad / groups;
}

// Which (row tile, query group) items worker `w` of `nw` (a CTA, or a CTA pair) processes; item = tile *
// n_groups + group: round-robin, so that at any moment the workers are on the same few row tiles (one HBM
// read per tile, L2 hits for its other query groups; working set ~5 tiles).  Measured alternative (ncu r01h):
// one contiguous item range per worker -- the re-use then is within a worker, but 74 different row tiles plus
// the query groups are live at once, the L2 hit rate falls from 93 % to 84 % and DRAM reads rise from 1.66x
// to 3.5x the corpus per pass.  (The 1.66x itself is drift: since the 8-warp epilogue the workers are no
// longer paced by a common bottleneck, and late ones find their tile evicted; harmless at 8 % DRAM load.)
__device__ __forceinline__ void item_range(int n_items, int n_groups, int w, int nw, int& begin, int& end,
                                           int& step) {
    (void)n_groups;
    step = nw;
    begin = w;
    end = (w < n_items) ? w + ((n_items - 1 - w) / nw + 1) * nw : w;
}

// store tile scanned as tile `tile` of this launch (see GemmParams::perm_mul)
__device__ __forceinline__ int64_t scan_tile(const GemmParams& p, int tile) {
    const int64_t g = p.tile_begin + tile;
    return p.perm_mul ? (g * p.perm_mul) % p.perm_tiles : g;
}

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    // K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO unused,
    // descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// Append the survivors a warp staged in shared memory to their queries' pools: every lane takes
// entries, so 32-64 global atomics are in flight per round trip instead of one per admitted column.
__device__ __forceinline__ void flush_staged(const uint2* stg, int wcount, int lane, int64_t row_warp0,
                                             const GemmParams& p) {
    __syncwarp();
    if (wcount >= kAggregateFrom) {
        // Many survivors = a slab whose threshold is still loose (5 % of all scores pass right after the first
        // slab): entries of one query sit next to each other in the staging buffer (they were staged column by
        // column), so one atomic per RUN of equal queries within the 32 entries a warp takes at a time reserves
        // the slots of the whole run.
        const uint32_t lanemask_lt = (1u << lane) - 1u;
        for (int e0 = 0; e0 < wcount; e0 += 64) {  // two chunks of 32 per round: as many atomics in flight as below
            const int ea = e0 + lane, eb = e0 + 32 + lane;
            const bool va = ea < wcount, vb = eb < wcount;
            const uint2 na = va ? stg[ea] : make_uint2(0u, 0u);
            const uint2 nb = vb ? stg[eb] : make_uint2(0u, 0u);
            const int qa = va ? (int)(na.y >> 5) : (0x40000000 | lane);
            const int qb = vb ? (int)(nb.y >> 5) : (0x40000000 | lane);
            const uint32_t sa = __match_any_sync(0xffffffffu, qa);
            const uint32_t sb = __match_any_sync(0xffffffffu, qb);
            const int la = __ffs(sa) - 1, lb = __ffs(sb) - 1;
            int ba = 0, bb = 0;
            if (va && lane == la) ba = atomicAdd(p.pool_cnt + qa, __popc(sa));
            if (vb && lane == lb) bb = atomicAdd(p.pool_cnt + qb, __popc(sb));
            ba = __shfl_sync(0xffffffffu, ba, la);
            bb = __shfl_sync(0xffffffffu, bb, lb);
            const int pa = ba + __popc(sa & lanemask_lt), pb = bb + __popc(sb & lanemask_lt);
            if (va && pa < kPoolCap) {
                p.pool_scores[(size_t)qa * kPoolCap + pa] = __uint_as_float(na.x);
                p.pool_ids[(size_t)qa * kPoolCap + pa] = (int32_t)(row_warp0 + (int)(na.y & 31u));
            }
            if (vb && pb < kPoolCap) {
                p.pool_scores[(size_t)qb * kPoolCap + pb] = __uint_as_float(nb.x);
                p.pool_ids[(size_t)qb * kPoolCap + pb] = (int32_t)(row_warp0 + (int)(nb.y & 31u));
            }
        }
        __syncwarp();
        return;
    }
    for (int e = lane; e < wcount; e += 64) {
        const bool two = (e + 32 < wcount);
        const uint2 en0 = stg[e];
        const uint2 en1 = two ? stg[e + 32] : make_uint2(0u, 0u);
        const int qa = (int)(en0.y >> 5), qb = (int)(en1.y >> 5);
        const int pos0 = atomicAdd(p.pool_cnt + qa, 1);
        const int pos1 = two ? atomicAdd(p.pool_cnt + qb, 1) : kPoolCap;
        if (pos0 < kPoolCap) {
            p.pool_scores[(size_t)qa * kPoolCap + pos0] = __uint_as_float(en0.x);
            p.pool_ids[(size_t)qa * kPoolCap + pos0] = (int32_t)(row_warp0 + (int)(en0.y & 31u));
        }
        if (pos1 < kPoolCap) {
            p.pool_scores[(size_t)qb * kPoolCap + pos1] = __uint_as_float(en1.x);
            p.pool_ids[(size_t)qb * kPoolCap + pos1] = (int32_t)(row_warp0 + (int)(en1.y & 31u));
        }
    }
    __syncwarp();
}

// One epilogue warp, one finished accumulator: 32 TMEM lanes (corpus rows row_warp0 .. +31) x ncols
// query columns starting at TMEM address `taddr`.  Scores = accumulator x per-row multiplier (NaN for
// tombstoned / out-of-range rows).  Dense slab: every score goes to its own pool slot.  Otherwise the
// score is compared with the query's admission threshold (thresholds of 32 queries are fetched as one
// batch of broadcast, L1-resident loads while the TMEM load is in flight; padded queries carry +inf)
// and the rare survivors are staged in the warp's shared-memory buffer (ballot + popc, no atomics).
// `release()` is called once every tcgen05.ld of this accumulator has completed -- before the
// global-atomic flush, so the MMA warp gets the accumulator back as early as possible.
// `dense_slot0` = dense slab only: the slot of row_warp0 in the destination (row - row_begin, or the position in
// the sampled wide slab).  `nt_local` = TMEM columns this warp owns from `taddr` on (the whole query group, or half of it when two
// warps share a lane quarter), `ncols` <= nt_local of them are real queries; `stage_cap` = entries in `stg`.
template <typename Release>
__device__ __forceinline__ void epilogue_item(const GemmParams& p, uint32_t taddr, int64_t row_warp0,
                                              int64_t dense_slot0, int lane, int q0, int ncols, int nt_local,
                                              uint2* stg, int stage_cap, Release release) {
    const uint32_t lanemask_lt = (1u << lane) - 1u;
    const int64_t row = row_warp0 + lane;
    const bool row_ok = row < p.row_end;
    const float mul = row_ok ? __ldg(p.row_mul + row) : __int_as_float(0x7fc00000);
    int wcount = 0;  // survivors staged by this warp (warp-uniform)
    for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        const bool wide = (nt_local - c0 >= 32);
        if (wide) {
            ptx::tmem_ld_32x32(taddr + (uint32_t)c0, v);
        } else {
            uint32_t w[16];
            ptx::tmem_ld_32x16(taddr + (uint32_t)c0, w);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                v[j] = w[j];
                v[16 + j] = 0u;
            }
        }
        float4 t4[8];
        if (!p.dense) {
            const float4* tp = reinterpret_cast<const float4*>(p.pool_thr + q0 + c0);
#pragma unroll
            for (int g = 0; g < 8; ++g)
                t4[g] = (wide || g < 4) ? __ldg(tp + g) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        }
        ptx::tmem_ld_wait();
        if (p.dense) {
            const int cend = (ncols - c0 < 32) ? (ncols - c0) : 32;
            const size_t slot = (size_t)(dense_slot0 + lane);  // where lane 0's row goes in the dense destination
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (j < cend) {
                    // (a row beyond the end of the store -- the partial last tile can sit anywhere in a permuted
                    // slab -- gets -inf like a tombstoned one)
                    const float s = row_ok ? __uint_as_float(v[j]) * mul : -INFINITY;
                    const size_t pos = (size_t)(q0 + c0 + j) * (size_t)p.dense_stride + slot;
                    p.dense_scores[pos] = (s == s) ? s : -INFINITY;
                    p.dense_ids[pos] = row_ok ? (int32_t)row : 0;
                }
            }
        } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float th[4] = {t4[g].x, t4[g].y, t4[g].z, t4[g].w};
                float sc[4];
                bool pass[4];
                bool any = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    sc[u] = __uint_as_float(v[4 * g + u]) * mul;
                    pass[u] = sc[u] >= th[u];
                    any |= pass[u];
                }
                if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t m = __ballot_sync(0xffffffffu, pass[u]);
                        if (pass[u]) {
                            const int e = wcount + __popc(m & lanemask_lt);
                            stg[e] = make_uint2(__float_as_uint(sc[u]),
                                                ((uint32_t)(q0 + c0 + 4 * g + u) << 5) | (uint32_t)lane);
                        }
                        wcount += __popc(m);
                    }
                    if (wcount > stage_cap - 128) {
                        flush_staged(stg, wcount, lane, row_warp0, p);
                        wcount = 0;
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncwarp();
    release();
    if (wcount > 0) flush_staged(stg, wcount, lane, row_warp0, p);
}

}  // namespace cmw
