// pool.cu -- query preparation, candidate-pool maintenance (radix-select compaction; the merge kernel of the
// wide first slab that small batches take), K3 (exact fp64 rescoring + final deterministic selection with an
// exactness certificate) and K5 (cross-shard merge).
//
// Ordering contract everywhere: score descending, then id ascending (BASELINE.json north_star:
// "ties broken by lower id"); results are best-first like collection.query() of
// rag_engine/storage/vector_store.py:59-66.
#include "common.cuh"
#include "ptx.cuh"

namespace cmw {

// ---------------------------------------------------------------------------------------------
// query preparation: fp64 norm, scaled fp32 copy (K1), bf16 copy padded with zero rows (K2)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_queries_kernel(const float* __restrict__ q, int batch, int bpad, int dim, int metric,
                    double* __restrict__ qn64, double* __restrict__ q4, double* __restrict__ qres,
                    float* __restrict__ q_f32, __nv_bfloat16* __restrict__ q_bf16, int half_tiles,
                    float* __restrict__ q_tf32, Pool pool, int dense_count, Pool seg, int wide_rows) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // (this kernel overwrites the workspace the previous search's last kernels may still be reading)
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    if (b >= bpad) return;
    // wide first slab: the 16 scratch segments of this query start out holding their share of its rows
    if (seg.cnt != nullptr && b < batch && lane < kWideSegments) {
        int c = wide_rows - lane * kPoolCap;
        c = c < 0 ? 0 : (c > kPoolCap ? kPoolCap : c);
        seg.cnt[b * kWideSegments + lane] = c;
        seg.thr[b * kWideSegments + lane] = -INFINITY;
        seg.ovf[b * kWideSegments + lane] = 0;
    }
    const int nvec = dim >> 2;
    // pool state for the new search: the dense slab will fill `dense_count` slots of every real query;
    // cnt/thr/ovf hold bpad entries and the padded queries of the last K2 group keep thr = +inf so that
    // the epilogue needs no column mask
    if (lane == 0) {
        pool.cnt[b] = (b < batch) ? dense_count : 0;
        pool.thr[b] = (b < batch) ? -INFINITY : INFINITY;
        pool.ovf[b] = 0;
    }
    if (b >= batch) {
        if (q_bf16 != nullptr) {
            uint2* out = reinterpret_cast<uint2*>(q_bf16 + (size_t)b * dim);
            for (int c = lane; c < nvec; c += 32) out[c] = make_uint2(0u, 0u);
        }
        if (q_tf32 != nullptr) {
            float4* out = reinterpret_cast<float4*>(q_tf32 + (size_t)b * dim);
            for (int c = lane; c < nvec; c += 32) out[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    const float4* in = reinterpret_cast<const float4*>(q + (size_t)b * dim);
    double acc = 0.0;
    for (int c = lane; c < nvec; c += 32) {
        float4 v = __ldg(in + c);
        acc += (double)v.x * (double)v.x;
        acc += (double)v.y * (double)v.y;
        acc += (double)v.z * (double)v.z;
        acc += (double)v.w * (double)v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const double nrm = sqrt(acc);
    // the filter always sees the L2-normalised query (pool scores in filter units, see cert_qscale)
    (void)metric;
    const double scale = nrm > 0.0 ? 1.0 / nrm : 0.0;
    if (lane == 0) qn64[b] = nrm;
    {
        // |q/|q||_4 for the certificate bound
        const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
        double s4 = 0.0;
        for (int c = lane; c < nvec; c += 32) {
            float4 v = __ldg(in + c);
            const double x = (double)v.x * inv, y = (double)v.y * inv, z = (double)v.z * inv, w = (double)v.w * inv;
            s4 += x * x * x * x + y * y * y * y + z * z * z * z + w * w * w * w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s4 += __shfl_xor_sync(0xffffffffu, s4, o);
        if (lane == 0) q4[b] = sqrt(sqrt(s4));
    }
    float4* of = reinterpret_cast<float4*>(q_f32 + (size_t)b * dim);
    double res2 = 0.0;  // |scaled query - its bf16 tile|^2, from the values actually written
    for (int c = lane; c < nvec; c += 32) {
        float4 v = __ldg(in + c);
        const double ex = (double)v.x * scale, ey = (double)v.y * scale, ez = (double)v.z * scale,
                     ew = (double)v.w * scale;
        float4 w = make_float4((float)ex, (float)ey, (float)ez, (float)ew);
        of[c] = w;
        if (q_bf16 != nullptr) {
            uint2* ob = reinterpret_cast<uint2*>(q_bf16 + (size_t)b * dim);
            float sx, sy, sz, sw;
            const uint32_t tx = to_tile16(w.x, half_tiles, sx), ty = to_tile16(w.y, half_tiles, sy);
            const uint32_t tz = to_tile16(w.z, half_tiles, sz), tw = to_tile16(w.w, half_tiles, sw);
            ob[c] = make_uint2(tx | (ty << 16), tz | (tw << 16));
            const double dx = ex - (double)sx, dy = ey - (double)sy, dz = ez - (double)sz, dw = ew - (double)sw;
            res2 += dx * dx + dy * dy + dz * dz + dw * dw;
        }
        if (q_tf32 != nullptr) {
            // round to nearest even at 10 mantissa bits: exactly representable, so the MMA narrows nothing further
            auto rn = [](float x) {
                uint32_t u = __float_as_uint(x);
                u += 0xfffu + ((u >> 13) & 1u);
                return __uint_as_float(u & 0xffffe000u);
            };
            const float4 t = make_float4(rn(w.x), rn(w.y), rn(w.z), rn(w.w));
            reinterpret_cast<float4*>(q_tf32 + (size_t)b * dim)[c] = t;
            const double dx = ex - (double)t.x, dy = ey - (double)t.y, dz = ez - (double)t.z, dw = ew - (double)t.w;
            res2 += dx * dx + dy * dy + dz * dz + dw * dw;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) res2 += __shfl_xor_sync(0xffffffffu, res2, o);
    if (lane == 0) qres[b] = sqrt(res2) * (1.0 + 1e-9) + 1e-12;  // the normalised query has norm 1; rounded up a little
}

int launch_prep_queries(const float* q, int batch, int bpad, int dim, int metric, double* qn64, double* q4,
                        double* qres, float* q_f32, __nv_bfloat16* q_bf16, int half_tiles, float* q_tf32, Pool pool,
                        int dense_count, Pool seg, int wide_rows, cudaStream_t stream) {
    const int wpb = 8;
    CMW_CUDA_OK(launch_kernel(prep_queries_kernel, dim3((bpad + wpb - 1) / wpb), dim3(wpb * 32), 0, stream,
                              batch <= kWideDenseMaxBatch, q, batch, bpad, dim, metric, qn64, q4, qres, q_f32, q_bf16,
                              half_tiles, q_tf32, pool, dense_count, seg, wide_rows));
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// pool maintenance
// ---------------------------------------------------------------------------------------------
// One CTA per query.  Intermediate slabs only need the best kprime entries and the admission
// threshold, not their order: an MSB-first radix select (4 passes of 8 bits over the order-preserving
// score keys, shared-memory histograms) finds the kprime-th best score T, entries with score >= T are
// compacted to the front (all ties at T are kept, so the `score >= thr` admission rule and the pool
// stay consistent) and thr is raised to T.  The last call (final = 1) also sorts the survivors by
// (score desc, id asc) and truncates to kprime -- the order K3 / the bf16 emit rely on.
//
// Entries whose score is -inf are ABSENT: that is what the dense first slab writes for tombstoned rows and
// for the slots past the end of the store.  They never count towards kprime and never survive, so a slab of
// dead rows leaves the pool empty and thr at -inf instead of filling the pool with 4096 placeholders (which
// made every later admission fall off the end of the pool).
constexpr int kCompactThreads = 256;
constexpr int kCompactPer = kPoolCap / kCompactThreads;  // 16 entries per thread, in registers

struct SelectShared {
    int hist[256];
    int warp_sums[kCompactThreads / 32];
    int sel_digit, sel_below;
};

// key of one pool entry (ascending = better); returns false for an absent entry
__device__ __forceinline__ bool load_entry(float score, int32_t row, uint32_t& key, int32_t& rid) {
    key = 0xffffffffu;
    rid = -1;
    if (!(score > -INFINITY)) return false;  // -inf and NaN: absent
    key = (uint32_t)(desc_key(score, 0u) >> 32);
    rid = row;
    return true;
}

// exclusive prefix of `mine` over the CTA and the CTA-wide total (two barriers)
__device__ __forceinline__ int block_exclusive_scan(int mine, int& total, SelectShared& sm) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    int v = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    __syncthreads();  // warp_sums may still be read by the previous use
    if (lane == 31) sm.warp_sums[w] = v;
    __syncthreads();
    int base = 0;
    total = 0;
#pragma unroll
    for (int ww = 0; ww < kCompactThreads / 32; ++ww) {
        const int x = sm.warp_sums[ww];
        if (ww < w) base += x;
        total += x;
    }
    return base + v - mine;
}

// The selection shared by the compaction kernel and the merge kernel of the wide first slab.  Every thread
// holds kCompactPer entries in registers (rid < 0 = absent).  Keeps the best kprime (and every tie at the
// cut), writes them to dst_sc / dst_id -- unordered, or for final != 0 sorted by (score desc, id asc) and
// truncated to kprime -- and publishes cnt / thr / ovf of pool slot `b`.
__device__ __forceinline__ void select_and_compact(uint32_t (&key)[kCompactPer], int32_t (&rid)[kCompactPer],
                                                   int kprime, int final, bool overflowed, Pool pool, int b,
                                                   uint64_t* sort_keys, SelectShared& sm) {
    const int t = threadIdx.x;
    float* sc = pool.scores + (size_t)b * kPoolCap;
    int32_t* id = pool.ids + (size_t)b * kPoolCap;
    int valid_mine = 0;
#pragma unroll
    for (int j = 0; j < kCompactPer; ++j) valid_mine += rid[j] >= 0 ? 1 : 0;
    int valid;
    block_exclusive_scan(valid_mine, valid, sm);
    uint32_t T = 0xffffffffu;  // keep everything that is present
    const bool select = valid > kprime;
    if (select) {
        // The keys of one pool are scores from a narrow range: their leading bits (sign, exponent, often a few
        // mantissa bits) are all the same.  Find that common prefix first (block-wide AND / OR) and run the radix
        // passes only over the bits below it: usually 3 passes of 8 bits instead of 4, and no pass in which all
        // 4096 shared-memory atomics land on one or two bins.
        uint32_t k_and = 0xffffffffu, k_or = 0u;
#pragma unroll
        for (int j = 0; j < kCompactPer; ++j) {
            if (rid[j] >= 0) {
                k_and &= key[j];
                k_or |= key[j];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            k_and &= __shfl_xor_sync(0xffffffffu, k_and, o);
            k_or |= __shfl_xor_sync(0xffffffffu, k_or, o);
        }
        __syncthreads();  // hist is free (nothing has used it yet in this call; a previous pass may have)
        if ((t & 31) == 0) {
            sm.hist[t >> 5] = (int)k_and;
            sm.hist[32 + (t >> 5)] = (int)k_or;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kCompactThreads / 32; ++w) {
            k_and &= (uint32_t)sm.hist[w];
            k_or |= (uint32_t)sm.hist[32 + w];
        }
        __syncthreads();
        const uint32_t diff = k_and ^ k_or;
        const int common = diff ? __clz(diff) : 32;  // leading bits shared by every present key
        uint32_t mask = common ? ~(0xffffffffu >> common) : 0u;
        if (common == 32) mask = 0xffffffffu;
        uint32_t prefix = k_and & mask;
        int remaining = kprime;
        for (int top = 32 - common; top > 0; top -= 8) {
            // digit = bits [shift, shift + 8); the last digit may overlap bits that are already decided, which
            // are equal for every key still in play
            const int shift = top >= 8 ? top - 8 : 0;
            sm.hist[t] = 0;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kCompactPer; ++j) {
                if (rid[j] >= 0 && (key[j] & mask) == prefix) atomicAdd(&sm.hist[(key[j] >> shift) & 255u], 1);
            }
            __syncthreads();
            // inclusive scan of the 256 bins (one per thread)
            const int h = sm.hist[t];
            int total;
            const int cum = block_exclusive_scan(h, total, sm) + h;
            if (cum >= remaining && cum - h < remaining) {
                sm.sel_digit = t;
                sm.sel_below = cum - h;
            }
            __syncthreads();
            prefix = (prefix & ~(0xffu << shift)) | ((uint32_t)sm.sel_digit << shift);
            mask |= 0xffu << shift;
            remaining -= sm.sel_below;
        }
        T = prefix;
    }
    // stream compaction of the kept entries (every entry is in registers)
    int mine = 0;
#pragma unroll
    for (int j = 0; j < kCompactPer; ++j) mine += (rid[j] >= 0 && key[j] <= T) ? 1 : 0;
    int total;
    int pos = block_exclusive_scan(mine, total, sm);
    if (!final) {
#pragma unroll
        for (int j = 0; j < kCompactPer; ++j) {
            if (rid[j] >= 0 && key[j] <= T) {
                sc[pos] = f32_from_orderable(~key[j]);
                id[pos] = rid[j];
                ++pos;
            }
        }
        if (t == 0) {
            pool.cnt[b] = total;
            if (select) pool.thr[b] = f32_from_orderable(~T);
            if (overflowed) pool.ovf[b] = 1;
        }
        return;
    }
    // final: sort the survivors by (score desc, id asc), keep kprime
    const int m = next_pow2(total < 2 ? 2 : total);
    for (int i = total + t; i < m; i += kCompactThreads) sort_keys[i] = ~0ull;
#pragma unroll
    for (int j = 0; j < kCompactPer; ++j) {
        if (rid[j] >= 0 && key[j] <= T) {
            sort_keys[pos] = ((uint64_t)key[j] << 32) | (uint64_t)(uint32_t)rid[j];
            ++pos;
        }
    }
    bitonic_sort_u64(sort_keys, m);
    const int keep = total < kprime ? total : kprime;
    for (int i = t; i < keep; i += kCompactThreads) {
        const uint64_t k = sort_keys[i];
        sc[i] = desc_key_score(k);
        id[i] = (int32_t)(uint32_t)k;
    }
    if (t == 0) {
        pool.cnt[b] = keep;
        if (total >= kprime) pool.thr[b] = desc_key_score(sort_keys[kprime - 1]);
        if (overflowed) pool.ovf[b] = 1;
    }
}

__global__ void __launch_bounds__(kCompactThreads, 5) pool_compact_kernel(Pool pool, int kprime, int final) {
    extern __shared__ __align__(16) uint8_t cmp_smem[];  // 32 KB: the staged entries, later the sort keys
    __shared__ SelectShared sm;
    __shared__ __align__(8) uint64_t stage_bar;
    const int b = blockIdx.x;
    const int t = threadIdx.x;
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int n_in = pool.cnt[b];
    const int n = n_in < kPoolCap ? n_in : kPoolCap;
    const float* sc = pool.scores + (size_t)b * kPoolCap;
    const int32_t* id = pool.ids + (size_t)b * kPoolCap;
    // The n entries come in through TWO bulk async copies (scores, ids: cp.async.bulk + mbarrier transaction count,
    // SASS UBLKCP) instead of 32 predicated loads per thread: with 48 registers per thread the compiler cannot keep
    // more than a few of those loads in flight, and the load phase was 60 % of the kernel's stall samples (ncu
    // source page, prof_r02_tail) -- a chain of L2 / HBM latencies per CTA.
    float* s_sc = reinterpret_cast<float*>(cmp_smem);
    int32_t* s_id = reinterpret_cast<int32_t*>(cmp_smem + (size_t)kPoolCap * sizeof(float));
    if (t == 0) {
        ptx::mbar_init(&stage_bar, 1);
        ptx::fence_barrier_init();
    }
    __syncthreads();
    if (n > 0) {
        if (t == 0) {
            const uint32_t bytes = ((uint32_t)n * 4u + 15u) & ~15u;  // rounds up inside the pool's own 16 KB row
            const uint64_t pol = ptx::l2_policy_evict_first();
            ptx::mbar_arrive_expect_tx(&stage_bar, 2u * bytes);
            ptx::bulk_g2s(s_sc, sc, bytes, &stage_bar, pol);
            ptx::bulk_g2s(s_id, id, bytes, &stage_bar, pol);
        }
        ptx::mbar_wait(&stage_bar, 0);
    }
    uint32_t key[kCompactPer];
    int32_t rid[kCompactPer];
#pragma unroll
    for (int j = 0; j < kCompactPer; ++j) {
        const int i = t + j * kCompactThreads;
        key[j] = 0xffffffffu;
        rid[j] = -1;
        if (i < n) load_entry(s_sc[i], s_id[i], key[j], rid[j]);
    }
    // (the barriers of the first block scan inside separate these reads from the in-place write-back and from the
    // sort keys that later take the staging buffer's place)
    select_and_compact(key, rid, kprime, final, n_in > kPoolCap, pool, b, reinterpret_cast<uint64_t*>(cmp_smem), sm);
}

int launch_pool_compact(Pool pool, int batch, int kprime, int final, cudaStream_t stream) {
    const size_t smem = (size_t)kPoolCap * sizeof(uint64_t);  // staging (scores | ids) = the sort keys of the final call
    CMW_CUDA_OK(launch_kernel(pool_compact_kernel, dim3(batch), dim3(kCompactThreads), smem, stream,
                              batch <= kWideDenseMaxBatch * kWideSegments, pool, kprime, final));
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Wide first slab (small batches): the filter wrote the scores and ids of up to 65536 rows into a scratch
// matrix laid out as 16 pool-sized segments per query.  Level 1 = the ordinary compaction kernel over the
// batch * 16 segments (each keeps its best kprime and ties, in parallel on as many SMs); level 2 = this
// kernel: one CTA per query gathers the segments' survivors (<= 16 * kprime + ties), runs the same selection
// on them, moves the kept entries into the query's pool and publishes thr.  An entry that is among the best
// kprime overall is among the best kprime of its segment, so the cut is exact; ties at the cut survive level
// 1 for the same reason.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCompactThreads, 5) wide_merge_kernel(Pool seg, Pool pool, int kprime, int final) {
    extern __shared__ __align__(16) uint8_t wm_smem[];  // 32 KB: the staged survivors, later the sort keys
    __shared__ SelectShared sm;
    __shared__ int seg_off[kWideSegments + 1];   // prefix sums of the segments' survivor counts
    __shared__ int seg_poff[kWideSegments + 1];  // the same with every count rounded up to 4 entries (16 bytes)
    __shared__ int seg_flag;
    __shared__ __align__(8) uint64_t stage_bar;
    const int b = blockIdx.x;
    const int t = threadIdx.x;
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    if (t < 32) {
        const int c = (t < kWideSegments) ? seg.cnt[b * kWideSegments + t] : 0;
        const int of = (t < kWideSegments) ? seg.ovf[b * kWideSegments + t] : 0;
        int v = c, vp = (c + 3) & ~3;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            const int up = __shfl_up_sync(0xffffffffu, vp, o);
            if (t >= o) {
                v += u;
                vp += up;
            }
        }
        if (t < kWideSegments) {
            seg_off[t + 1] = v;
            seg_poff[t + 1] = vp;
        }
        if (t == 0) {
            seg_off[0] = 0;
            seg_poff[0] = 0;
            ptx::mbar_init(&stage_bar, 1);
            ptx::fence_barrier_init();
        }
        const unsigned any = __ballot_sync(0xffffffffu, of != 0);
        if (t == 0) seg_flag = any != 0u;
    }
    __syncthreads();
    const int n_in = seg_off[kWideSegments];
    const int n_pad = seg_poff[kWideSegments];
    uint32_t key[kCompactPer];
    int32_t rid[kCompactPer];
    if (n_in <= kPoolCap && n_pad <= kPoolCap) {
        // The survivors of the 16 segments come in through bulk async copies (one pair per segment, 16-byte
        // granules: each segment's piece is padded to 4 entries in the staging buffer and the padding ignored)
        // instead of predicated loads with a 16-way segment search per entry -- the same latency chain the
        // compaction kernel had.
        float* s_sc = reinterpret_cast<float*>(wm_smem);
        int32_t* s_id = reinterpret_cast<int32_t*>(wm_smem + (size_t)kPoolCap * sizeof(float));
        if (n_pad > 0) {
            if (t == 0) {
                const uint64_t pol = ptx::l2_policy_evict_first();
                ptx::mbar_arrive_expect_tx(&stage_bar, 2u * (uint32_t)n_pad * 4u);
                for (int sgm = 0; sgm < kWideSegments; ++sgm) {
                    const uint32_t bytes = (uint32_t)(seg_poff[sgm + 1] - seg_poff[sgm]) * 4u;
                    if (bytes == 0) continue;
                    const size_t src = ((size_t)b * kWideSegments + sgm) * kPoolCap;
                    ptx::bulk_g2s(s_sc + seg_poff[sgm], seg.scores + src, bytes, &stage_bar, pol);
                    ptx::bulk_g2s(s_id + seg_poff[sgm], seg.ids + src, bytes, &stage_bar, pol);
                }
            }
            ptx::mbar_wait(&stage_bar, 0);
        }
#pragma unroll
        for (int j = 0; j < kCompactPer; ++j) {
            const int i = t + j * kCompactThreads;
            key[j] = 0xffffffffu;
            rid[j] = -1;
            if (i < n_pad) {
                int sgm = 0;
#pragma unroll
                for (int q = 1; q < kWideSegments; ++q) sgm += (i >= seg_poff[q]) ? 1 : 0;
                if (i - seg_poff[sgm] < seg_off[sgm + 1] - seg_off[sgm]) load_entry(s_sc[i], s_id[i], key[j], rid[j]);
            }
        }
    } else {
        // more survivors than one pool holds (ties, overflowed segments): the first kPoolCap, flagged below
        const int n = n_in < kPoolCap ? n_in : kPoolCap;
#pragma unroll
        for (int j = 0; j < kCompactPer; ++j) {
            const int i = t + j * kCompactThreads;
            key[j] = 0xffffffffu;
            rid[j] = -1;
            if (i < n) {
                int sgm = 0;
#pragma unroll
                for (int q = 1; q < kWideSegments; ++q) sgm += (i >= seg_off[q]) ? 1 : 0;
                const size_t src = ((size_t)b * kWideSegments + sgm) * kPoolCap + (size_t)(i - seg_off[sgm]);
                load_entry(seg.scores[src], seg.ids[src], key[j], rid[j]);
            }
        }
    }
    select_and_compact(key, rid, kprime, final, n_in > kPoolCap || seg_flag != 0, pool, b,
                       reinterpret_cast<uint64_t*>(wm_smem), sm);
}

int launch_wide_select(Pool seg, Pool pool, int batch, int kprime, int final, cudaStream_t stream) {
    int rc = launch_pool_compact(seg, batch * kWideSegments, kprime, 0, stream);
    if (rc) return rc;
    const size_t smem = (size_t)kPoolCap * sizeof(uint64_t);  // staging (scores | ids) = the sort keys of a final call
    CMW_CUDA_OK(launch_kernel(wide_merge_kernel, dim3(batch), dim3(kCompactThreads), smem, stream, true, seg, pool,
                              kprime, final));
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K3a: exact rescoring.  One warp per candidate: fp64 products accumulated in a fixed order
// (lane-sequential over the row, then an xor butterfly), so bit-identical rows get bit-identical
// scores and the result does not depend on which filter kernel produced the candidate.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rescore_kernel(const float* __restrict__ f32, const double* __restrict__ norm64,
               const float* __restrict__ live, int dim, Pool pool, int kprime, int metric,
               const float* __restrict__ q_raw, const CertParams cert, int k, double* __restrict__ exact) {
    const double* __restrict__ qn64 = cert.qn64;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int n = pool.cnt[b] < kprime ? pool.cnt[b] : kprime;
    // The pool is sorted by filter score.  With a_k its k-th filter score, every row whose filter score
    // is below a_k - 2*eps has an exact score below a_k - eps <= (k-th exact score): it cannot be in the
    // exact top-k, so it is not rescored (its slot gets -inf and sorts last).
    // Row shards: the k-th best filter score over ALL shards is handed in (cert.global_kth); the k rows behind it
    // have exact scores >= that - eps wherever they live, so the same cut holds against the global k-th exact score.
    double cut = -INFINITY;
    if (cert.global_kth != nullptr) cut = (double)cert.global_kth[b] - 2.0 * cert_eps(cert, b);
    else if (n > k) cut = (double)pool.scores[(size_t)b * kPoolCap + (k - 1)] - 2.0 * cert_eps(cert, b);
    const float4* qv = reinterpret_cast<const float4*>(q_raw + (size_t)b * dim);
    const int nvec = dim >> 2;
    // gridDim.x blocks share a query (many for small batches, one for large ones); a warp walks its
    // candidates in pool order, several rows' loads in flight
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += gridDim.x * wpb) {
        if ((double)pool.scores[(size_t)b * kPoolCap + j] < cut) {
            if (lane == 0) exact[(size_t)b * kprime + j] = -INFINITY;
            continue;
        }
        const int32_t id = pool.ids[(size_t)b * kPoolCap + j];
        const float4* row = reinterpret_cast<const float4*>(f32 + (size_t)id * dim);
        double acc = 0.0;
        if (nvec == 384) {
            // D = 1536: the whole row (12 x 128-bit per lane) is requested before the first use, so a
            // warp keeps 6 KB in flight; same summation order as the generic loop
            float4 v[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) v[i] = __ldg(row + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float4 w = __ldg(qv + lane + 32 * i);
                acc += (double)v[i].x * (double)w.x;
                acc += (double)v[i].y * (double)w.y;
                acc += (double)v[i].z * (double)w.z;
                acc += (double)v[i].w * (double)w.w;
            }
        } else {
#pragma unroll 4
            for (int c = lane; c < nvec; c += 32) {
                const float4 v = __ldg(row + c);
                const float4 w = __ldg(qv + c);
                acc += (double)v.x * (double)w.x;
                acc += (double)v.y * (double)w.y;
                acc += (double)v.z * (double)w.z;
                acc += (double)v.w * (double)w.w;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            double s;
            const float lv = live[id];
            if (lv != lv) {
                s = -INFINITY;  // tombstoned row that entered through the dense slab
            } else if (metric == CMW_METRIC_COSINE) {
                const double den = qn64[b] * norm64[id];
                s = den > 0.0 ? acc / den : 0.0;
            } else {
                s = acc;
            }
            exact[(size_t)b * kprime + j] = s;
        }
    }
}

// K3b: per query, sort the rescored candidates by (exact desc, id asc), emit k, certify.
__global__ void __launch_bounds__(256)
select_kernel(Pool pool, int k, int kprime, const double* __restrict__ exact, const CertParams cert,
              int64_t id_offset, float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
              double* __restrict__ out_scores64, int32_t* __restrict__ out_flags, double* __restrict__ out_aux) {
    extern __shared__ __align__(16) uint8_t sel_smem[];
    const int b = blockIdx.x;
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int n = pool.cnt[b] < kprime ? pool.cnt[b] : kprime;
    const int m = next_pow2(n < 2 ? 2 : n);
    uint64_t* hi = reinterpret_cast<uint64_t*>(sel_smem);
    uint64_t* lo = hi + m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        if (i < n) {
            hi[i] = ~f64_orderable(exact[(size_t)b * kprime + i]);
            lo[i] = (uint64_t)(uint32_t)pool.ids[(size_t)b * kPoolCap + i];
        } else {
            hi[i] = ~0ull;
            lo[i] = ~0ull;
        }
    }
    bitonic_sort_u128(hi, lo, m);
    const uint64_t neg_inf_key = ~f64_orderable(-INFINITY);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const bool ok = (j < n) && (hi[j] < neg_inf_key);
        double s = -INFINITY;
        int64_t id = -1;
        if (ok) {
            s = f64_from_orderable(~hi[j]);
            id = (int64_t)lo[j] + id_offset;
        }
        if (out_scores != nullptr) out_scores[(size_t)b * k + j] = (float)s;
        out_ids[(size_t)b * k + j] = id;
        if (out_scores64 != nullptr) out_scores64[(size_t)b * k + j] = s;
    }
    if (threadIdx.x == 0 && out_aux != nullptr) {
        // row shards: what the cross-shard certificate needs from this shard -- every row outside this pool has
        // filter score <= t (none, if the pool never filled), and eps with this shard's own residual bound
        // (both in exact-score units, so that shards may be compared)
        const double qs = cert_qscale(cert, b);
        out_aux[2 * b] = (n >= kprime) ? (double)pool.thr[b] * qs : -INFINITY;
        out_aux[2 * b + 1] = cert_eps(cert, b) * qs;
    }
    if (threadIdx.x == 0 && out_flags != nullptr) {
        int flag = 0;
        if (pool.ovf[b]) flag = CMW_FLAG_UNCERTIFIED;
        // (a shard of a row-sharded search cannot certify anything by itself: cmw_shard_merge does)
        if (n >= kprime && cert.global_kth == nullptr) {
            // rows outside the pool have filter score <= t, hence exact score <= t + eps
            const double t = (double)pool.thr[b];
            const double e = cert_eps(cert, b);
            const int kk = (k <= n) ? k : n;
            double kth = -INFINITY;
            if (kk >= 1 && hi[kk - 1] < neg_inf_key) {
                kth = f64_from_orderable(~hi[kk - 1]);
            }
            if (!(kth > (t + e) * cert_qscale(cert, b))) flag = CMW_FLAG_UNCERTIFIED;  // t, e: filter units
        }
        out_flags[b] = flag;
    }
}

int launch_rescore_select(const Store* s, Pool pool, int batch, int k, int kprime, int metric,
                          const float* q_raw, const CertParams& cert, double* exact_ws,
                          float* out_scores, int64_t* out_ids, double* out_scores64,
                          int32_t* out_flags, double* out_aux, cudaStream_t stream) {
    CMW_REQUIRE(s->f32 != nullptr, "CMW_MODE_F32_EXACT needs a store created with CMW_STORE_F32");
    const int wpb = 8;
    // enough blocks to fill the GPU a few times over, at most one warp per candidate
    int per_query = (s->sm_count * 16 + batch - 1) / batch;
    const int max_per_query = (kprime + wpb - 1) / wpb;
    if (per_query > max_per_query) per_query = max_per_query;
    if (per_query < 1) per_query = 1;
    dim3 grid(per_query, batch);
    const bool pdl = batch <= kWideDenseMaxBatch;
    CMW_CUDA_OK(launch_kernel(rescore_kernel, grid, dim3(wpb * 32), 0, stream, pdl, s->f32, s->norm64, s->live, s->dim,
                              pool, kprime, metric, q_raw, cert, k, exact_ws));
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    const size_t smem = (size_t)next_pow2_host(kprime) * 16;
    CMW_CUDA_OK(launch_kernel(select_kernel, dim3(batch), dim3(256), smem, stream, pdl, pool, k, kprime, exact_ws, cert,
                              s->id_offset, out_scores, out_ids, out_scores64, out_flags, out_aux));
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

// bf16 mode: the pool is already sorted by the last compaction; emit its best k.
__global__ void pool_emit_kernel(Pool pool, int k, int64_t id_offset, const double* __restrict__ qn64, int metric,
                                 float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                                 double* __restrict__ out_scores64, int32_t* __restrict__ out_flags,
                                 double* __restrict__ out_aux) {
    const int b = blockIdx.x;
    const int n = pool.cnt[b];
    const float qs = metric == CMW_METRIC_IP ? (float)qn64[b] : 1.0f;  // filter units -> inner product
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        float s = -INFINITY;
        int64_t id = -1;
        if (j < n) {
            const float v = pool.scores[(size_t)b * kPoolCap + j];
            if (v > -INFINITY) {
                s = v * qs;
                id = (int64_t)pool.ids[(size_t)b * kPoolCap + j] + id_offset;
            }
        }
        if (out_scores != nullptr) out_scores[(size_t)b * k + j] = s;
        out_ids[(size_t)b * k + j] = id;
        if (out_scores64 != nullptr) out_scores64[(size_t)b * k + j] = (double)s;
    }
    if (threadIdx.x == 0 && out_flags != nullptr) out_flags[b] = pool.ovf[b] ? CMW_FLAG_UNCERTIFIED : 0;
    if (threadIdx.x == 0 && out_aux != nullptr) {  // approximate mode: nothing to certify across shards
        out_aux[2 * b] = -INFINITY;
        out_aux[2 * b + 1] = 0.0;
    }
}

int launch_pool_emit(const Store* s, Pool pool, int batch, int k, const double* qn64, int metric, float* out_scores,
                     int64_t* out_ids, double* out_scores64, int32_t* out_flags, double* out_aux,
                     cudaStream_t stream) {
    pool_emit_kernel<<<batch, 128, 0, stream>>>(pool, k, s->id_offset, qn64, metric, out_scores, out_ids,
                                               out_scores64, out_flags, out_aux);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K5: cross-shard merge after the all-gather: G lists of k_in per query -> k_out.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
merge_kernel(const double* __restrict__ scores, const int64_t* __restrict__ ids, int G, int B,
             int k_in, int k_out, float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
             double* __restrict__ out_scores64) {
    extern __shared__ __align__(16) uint8_t mrg_smem[];
    const int b = blockIdx.x;
    const int n = G * k_in;
    const int m = next_pow2(n < 2 ? 2 : n);
    uint64_t* hi = reinterpret_cast<uint64_t*>(mrg_smem);
    uint64_t* lo = hi + m;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        uint64_t h = ~0ull, l = ~0ull;
        if (i < n) {
            const int g = i / k_in, j = i - g * k_in;
            const size_t src = ((size_t)g * B + b) * k_in + j;
            const int64_t id = ids[src];
            const double s = scores[src];
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
// Row-sharded search: the three small kernels either side of the two all-gathers.
// ---------------------------------------------------------------------------------------------
// after the filter half: the best k FILTER scores of every pool (sorted by the final compaction), -inf padded
__global__ void pool_topk_scores_kernel(Pool pool, int k, float* __restrict__ out) {
    const int b = blockIdx.x;
    const int n = pool.cnt[b];
    for (int j = threadIdx.x; j < k; j += blockDim.x)
        out[(size_t)b * k + j] = j < n ? pool.scores[(size_t)b * kPoolCap + j] : -INFINITY;
}

int launch_pool_topk_scores(Pool pool, int batch, int k, float* out, cudaStream_t stream) {
    pool_topk_scores_kernel<<<batch, 128, 0, stream>>>(pool, k, out);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

// after all-gather 1: the k-th best filter score over all shards, per query (-inf when fewer than k exist).
// One WARP per query: MSB-first bitwise selection on the order-preserving keys -- 32 rounds of "how many keys have
// this prefix and a 1 in the next bit" over the G*k gathered scores (3 KB per query, L1-resident after the first
// round) -- instead of sorting them: at G = 8, k = 100 a sort of 1024 keys per query cost more than the exchange.
__global__ void __launch_bounds__(256)
shard_kth_kernel(const float* __restrict__ gathered, int G, int B, int k, float* __restrict__ out_kth) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int n = G * k;
    // ascending orderable key = ascending score; absent entries (-inf) and NaN never count
    auto key_of = [&](int i) -> uint32_t {
        const int g = i / k, j = i - g * k;
        const float v = gathered[((size_t)g * B + b) * k + j];
        return (v > -INFINITY) ? f32_orderable(v) : 0u;  // 0 is below every real score's key
    };
    int valid = 0;
    for (int i = lane; i < n; i += 32) valid += key_of(i) != 0u ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    if (valid < k) {
        if (lane == 0) out_kth[b] = -INFINITY;
        return;
    }
    uint32_t prefix = 0u;
    int want = k;  // the want-th LARGEST key among those matching the prefix so far
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t probe = prefix | (1u << bit);
        const uint32_t mask = ~((1u << bit) - 1u);
        int ones = 0;
        for (int i = lane; i < n; i += 32) ones += ((key_of(i) & mask) == probe) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ones += __shfl_xor_sync(0xffffffffu, ones, o);
        if (ones >= want) prefix = probe;  // the answer has this bit set
        else want -= ones;                 // it lies among the keys with a 0 here
    }
    if (lane == 0) out_kth[b] = f32_from_orderable(prefix);
}

// The same selection with the keys in REGISTERS (G*k <= 32 * KPL: 8 shards x top-100 is 25 keys per lane): every
// key is read from memory once (one integer division each), the 32 rounds are KPL compares per lane and one
// redux.sync.  The memory-resident kernel above took 0.22 ms per batch of 4096 at G = 8 -- a tenth of the whole
// 8-GPU step -- for 3 KB of input per query.
template <int KPL>
__global__ void __launch_bounds__(256)
shard_kth_reg_kernel(const float* __restrict__ gathered, int G, int B, int k, float* __restrict__ out_kth) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int n = G * k;
    uint32_t key[KPL];
    int valid = 0;
    uint32_t all_and = 0xffffffffu, all_or = 0u;
#pragma unroll
    for (int t = 0; t < KPL; ++t) {
        const int i = lane + 32 * t;
        uint32_t u = 0u;  // 0 is below every real score's key: absent entries (-inf), NaN and padding never count
        if (i < n) {
            const int g = i / k, j = i - g * k;
            const float v = __ldg(gathered + ((size_t)g * B + b) * k + j);
            if (v > -INFINITY) u = f32_orderable(v);
        }
        key[t] = u;
        if (u != 0u) {
            ++valid;
            all_and &= u;
            all_or |= u;
        }
    }
    valid = __reduce_add_sync(0xffffffffu, valid);
    if (valid < k) {
        if (lane == 0) out_kth[b] = -INFINITY;
        return;
    }
    all_and = __reduce_and_sync(0xffffffffu, all_and);
    all_or = __reduce_or_sync(0xffffffffu, all_or);
    // bits on which every valid key agrees need no round: they are the answer's bits too
    const uint32_t differ = all_and ^ all_or;
    uint32_t prefix = all_and & ~differ;
    uint32_t decided = ~differ;  // mask of the bits of `prefix` that are settled
    int want = k;  // the want-th LARGEST among the valid keys matching the settled bits
    for (int bit = 31; bit >= 0; --bit) {
        if (!((differ >> bit) & 1u)) continue;
        const uint32_t probe = prefix | (1u << bit);
        const uint32_t mask = decided | (1u << bit);
        // only bits ABOVE `bit` (all settled) and `bit` itself take part: lower settled bits are common to all
        const uint32_t hi_mask = mask & ~((1u << bit) - 1u);
        int ones = 0;
#pragma unroll
        for (int t = 0; t < KPL; ++t) ones += (key[t] != 0u && (key[t] & hi_mask) == (probe & hi_mask)) ? 1 : 0;
        ones = __reduce_add_sync(0xffffffffu, ones);
        if (ones >= want) prefix = probe;
        else want -= ones;
        decided |= (1u << bit);
    }
    if (lane == 0) out_kth[b] = f32_from_orderable(prefix);
}

// after all-gather 2: G blocks -> the global top k_out by (exact score desc, id asc) and the cross-shard
// certificate: every row outside shard g's pool has filter score <= t_g, hence exact score <= t_g + eps_g; rows
// inside a pool that were not rescored lie below the global cut (see rescore_kernel).  So the merged ids are the
// oracle's as soon as the k-th merged exact score exceeds max_g (t_g + eps_g).
__global__ void __launch_bounds__(256)
shard_merge_kernel(const uint8_t* __restrict__ blocks, int G, int B, int k, int k_out,
                   float* __restrict__ out_scores, int64_t* __restrict__ out_ids, double* __restrict__ out_scores64,
                   int32_t* __restrict__ out_flags, const int32_t* __restrict__ peer_status) {
    extern __shared__ __align__(16) uint8_t sm_smem[];
    __shared__ int warp_tot[8];
    const int b = blockIdx.x;
    if (peer_status != nullptr && *peer_status != 0) {
        // an exchange over peer memory did not complete (cmw_peer_gather): nothing below can be trusted
        for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
            out_scores[(size_t)b * k_out + j] = -INFINITY;
            out_ids[(size_t)b * k_out + j] = -1;
            if (out_scores64 != nullptr) out_scores64[(size_t)b * k_out + j] = -INFINITY;
        }
        if (threadIdx.x == 0 && out_flags != nullptr) out_flags[b] = *peer_status;
        return;
    }
    const int n_all = G * k;
    const int m_all = next_pow2(n_all < 2 ? 2 : n_all);
    uint64_t* hi = reinterpret_cast<uint64_t*>(sm_smem);
    uint64_t* lo = hi + m_all;
    const ShardBlock lay = shard_block(B, k);
    // With the global rescoring cut most of a shard's k slots are empty (only the candidates that can still reach
    // the global top-k were rescored): pack the valid ones first and sort next_pow2(valid) keys, not G*k.
    int base = 0;
    for (int i0 = 0; i0 < n_all; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        uint64_t h = ~0ull, l = ~0ull;
        bool ok = false;
        if (i < n_all) {
            const int g = i / k, j = i - g * k;
            const uint8_t* blk = blocks + (size_t)g * lay.total;
            const int64_t id = reinterpret_cast<const int64_t*>(blk + lay.ids)[(size_t)b * k + j];
            const double s = reinterpret_cast<const double*>(blk + lay.scores)[(size_t)b * k + j];
            if (id >= 0 && s == s && s > -INFINITY) {
                h = ~f64_orderable(s);
                l = (uint64_t)id;
                ok = true;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int before = 0, round_total = 0;
        for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) {
            if (ww < w) before += warp_tot[ww];
            round_total += warp_tot[ww];
        }
        if (ok) {
            const int pos = base + before + __popc(bal & ((1u << lane) - 1u));
            hi[pos] = h;
            lo[pos] = l;
        }
        base += round_total;
        __syncthreads();
    }
    const int n = base;  // valid candidates
    const int m = next_pow2(n < 2 ? 2 : n);
    for (int i = n + threadIdx.x; i < m; i += blockDim.x) {
        hi[i] = ~0ull;
        lo[i] = ~0ull;
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
    if (threadIdx.x == 0 && out_flags != nullptr) {
        int flag = 0;
        double bar = -INFINITY;  // max over shards of t_g + eps_g
        for (int g = 0; g < G; ++g) {
            const uint8_t* blk = blocks + (size_t)g * lay.total;
            flag |= reinterpret_cast<const int32_t*>(blk + lay.flags)[b];
            const double* aux = reinterpret_cast<const double*>(blk + lay.aux) + 2 * (size_t)b;
            if (aux[0] > -INFINITY && aux[0] + aux[1] > bar) bar = aux[0] + aux[1];
        }
        if (bar > -INFINITY) {
            // (a pool that filled holds at least K' >= k rows of which the k best were rescored: n >= k_out)
            const double kth = (n >= k_out) ? f64_from_orderable(~hi[k_out - 1]) : -INFINITY;
            if (!(kth > bar)) flag |= CMW_FLAG_UNCERTIFIED;
        }
        out_flags[b] = flag;
    }
}

}  // namespace cmw

using namespace cmw;

extern "C" int cmw_merge_topk(const double* scores_dev, const int64_t* ids_dev, int G, int B, int k_in,
                              int k_out, float* out_scores_dev, int64_t* out_ids_dev,
                              double* out_scores64_dev, void* stream) {
    CMW_REQUIRE(scores_dev && ids_dev && out_scores_dev && out_ids_dev, "cmw_merge_topk: NULL argument");
    CMW_REQUIRE(G >= 1 && B >= 0 && k_in >= 1 && k_out >= 1, "cmw_merge_topk: bad sizes");
    if (B == 0) return 0;
    const int n = G * k_in;
    CMW_REQUIRE(n <= 8192, "cmw_merge_topk: G*k_in = %d exceeds 8192", n);
    const size_t smem = (size_t)next_pow2_host(n) * 16;
    static SmemAttrCache smem_set;
    if (smem > 48 * 1024 && smem_set.needs(smem)) {
        CMW_CUDA_OK(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        smem_set.done(smem);
    }
    merge_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(scores_dev, ids_dev, G, B, k_in, k_out,
                                                        out_scores_dev, out_ids_dev, out_scores64_dev);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

namespace cmw {

int launch_shard_kth(const float* gathered, int G, int B, int k, float* out_kth, cudaStream_t stream) {
    const int n = G * k;
    CMW_REQUIRE(n <= 65536, "cmw_shard_kth: G*k = %d exceeds 65536", n);
    if (n <= 256) shard_kth_reg_kernel<8><<<(B + 7) / 8, 256, 0, stream>>>(gathered, G, B, k, out_kth);
    else if (n <= 1024) shard_kth_reg_kernel<32><<<(B + 7) / 8, 256, 0, stream>>>(gathered, G, B, k, out_kth);
    else shard_kth_kernel<<<(B + 7) / 8, 256, 0, stream>>>(gathered, G, B, k, out_kth);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_shard_merge(const void* blocks, int G, int B, int k, int k_out, float* out_scores, int64_t* out_ids,
                       double* out_scores64, int32_t* out_flags, const int32_t* peer_status, cudaStream_t stream) {
    const int n = G * k;
    CMW_REQUIRE(n <= 8192, "cmw_shard_merge: G*k = %d exceeds 8192", n);
    const size_t smem = (size_t)next_pow2_host(n < 2 ? 2 : n) * 16;
    static SmemAttrCache smem_set;
    if (smem > 48 * 1024 && smem_set.needs(smem)) {
        CMW_CUDA_OK(cudaFuncSetAttribute(shard_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set.done(smem);
    }
    shard_merge_kernel<<<B, 256, smem, stream>>>(reinterpret_cast<const uint8_t*>(blocks), G, B, k, k_out, out_scores,
                                                out_ids, out_scores64, out_flags, peer_status);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cmw
