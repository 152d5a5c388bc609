// gemm2.cu -- K2 for large batches: the CTA-pair (cta_group::2) variant of the tcgen05 filter.
//
// Same contract and epilogue as gemm.cu; what changes is the MMA shape and the operand traffic.  Two
// CTAs of a cluster (the two SMs of a TPC) cooperate on one item = 256 corpus rows x NT queries:
//   * each CTA TMA-loads its own 128 corpus rows (A) and HALF of the query tile (B, NT/2 rows), so the
//     query operand is fetched once per pair instead of once per SM: L2 -> shared-memory traffic per
//     k-block drops from 16 KB + NT*128 B to 16 KB + NT*64 B per SM;
//   * the leader CTA's MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = NT, K = 16): the tensor
//     cores of both SMs read A from their own shared memory and B from both halves; each SM's TMEM
//     receives the 128 x NT accumulator of its own rows;
//   * TMA completions of both CTAs signal the LEADER's full barrier (peer bit of the barrier address
//     cleared); tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers; the epilogue warps of
//     both CTAs release the accumulator on the leader's tmem_empty barrier (remote mbarrier arrive).
//   * 8 epilogue warps per CTA (two per TMEM lane quarter, 128 query columns each);
//   * work items come from cluster launch control (clusterlaunchcontrol.try_cancel): the grid holds one
//     pair per item, the resident pairs steal the pending ones -- a hardware dynamic scheduler that keeps
//     the pairs within one item of each other (L2 re-use of the corpus rows, one-item tails).
// Used when the padded batch is a multiple of 256 (NT = 256) and large enough to be tensor-bound; otherwise
// the 1-CTA kernel (HBM-bound regime) runs.
#include "gemm_common.cuh"

namespace cmw {

namespace ptx2 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(ptx::smem_u32(bar)),
        "r"(cta)
        : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit: address in the even (leader) CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* dst_smem, const CUtensorMap* m, int c0, int c1,
                                                uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(ptx::smem_u32(dst_smem)),
        "l"(m), "r"(c0), "r"(c1), "r"(ptx::smem_u32(bar) & kPeerBitMask), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     ptx::smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile(
        "{\n"
        ".reg .b16 msk;\n"
        "mov.b16 msk, 3;\n"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], msk;\n"
        "}\n" ::"r"(ptx::smem_u32(bar))
        : "memory");
}

// ---- cluster launch control (Blackwell's hardware work-stealing scheduler) -------------------------------
// The grid holds one cluster per work item; the clusters that are resident cancel the launch of pending
// ones and do their work themselves.  try_cancel writes a 16-byte response to the same shared-memory offset
// in every CTA of the cluster and completes 16 transaction bytes on each CTA's mbarrier at `bar`'s offset.
__device__ __forceinline__ void clc_try_cancel_multicast(void* resp_smem, uint64_t* bar) {
    asm volatile(
        "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 "
        "[%0], [%1];" ::"r"(ptx::smem_u32(resp_smem)),
        "r"(ptx::smem_u32(bar))
        : "memory");
}
// (valid, first CTA's blockIdx.x of the cancelled cluster); the trailing proxy fence orders this generic read
// before the async-proxy write of a later response into the same slot
__device__ __forceinline__ int clc_decode(const void* resp_smem) {
    uint32_t valid, x;
    asm volatile(
        "{\n"
        ".reg .pred p1;\n"
        ".reg .b128 r;\n"
        "ld.shared.b128 r, [%2];\n"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n"
        "selp.u32 %1, 1, 0, p1;\n"
        "mov.u32 %0, 0;\n"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %0, r;\n"
        "}\n"
        : "=r"(x), "=r"(valid)
        : "r"(ptx::smem_u32(resp_smem))
        : "memory");
    ptx::fence_proxy_async();
    return valid ? (int)x : -1;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint64_t* bar, uint32_t cta, uint32_t bytes) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [ra], %2;\n"
        "}\n" ::"r"(ptx::smem_u32(bar)),
        "r"(cta), "r"(bytes)
        : "memory");
}

}  // namespace ptx2

constexpr int kClcDepth = 4;  // responses in flight (the scheduler runs one item ahead; consumers lag <= 3)

constexpr int kGemm2Threads = 320;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int kGemm2EpiWarps = 8;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm_topk_kernel_2cta(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                      const GemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx2::cluster_ctarank();
    const bool leader = (rank == 0);

    uint8_t* stages = smem + ((1024u - (ptx::smem_u32(smem) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stages + (size_t)p.nstages * p.stage_bytes);
    uint64_t* full = bars;                          // [kMaxStages]  used in the leader only
    uint64_t* empty = bars + kMaxStages;            // [kMaxStages]  one set per CTA
    uint64_t* tmem_full = bars + 2 * kMaxStages;    // [2]           one set per CTA
    uint64_t* tmem_empty = tmem_full + 2;           // [2]           used in the leader only
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint2* stage_buf = reinterpret_cast<uint2*>(tmem_empty + 4);
    uint4* clc_resp = reinterpret_cast<uint4*>(stage_buf + (size_t)kGemm2EpiWarps * kStageCap2);  // [kClcDepth]
    uint64_t* clc_full = reinterpret_cast<uint64_t*>(clc_resp + kClcDepth);  // [kClcDepth] one set per CTA
    uint64_t* clc_empty = clc_full + kClcDepth;                              // [kClcDepth] used in the leader only

    const int n_items = p.n_tiles * p.n_groups;
    const int cid = blockIdx.x >> 1;
    const int ncl = gridDim.x >> 1;
    const int half_nt = p.nt >> 1;
    const int b_bytes = half_nt * kBlockK * 2;
    // Work distribution.  Static: pair c takes items c, c + ncl, ... (grid = one pair per two SMs).  Dynamic
    // (cluster launch control): the grid has one pair per item; a resident pair does its own item, then keeps
    // cancelling the launch of pending pairs and doing their items, until nothing is pending.  The pairs thus
    // stay within one item of each other whatever their individual speed: a row tile is read from HBM once
    // and found in L2 by its other query groups, and the tail of a slab is one item, not one static share.
    const bool dyn = p.dynamic != 0;
    // every consumer of the schedule (TMA lanes, MMA warp, epilogue warps of both CTAs) walks the same
    // sequence: `cons` counts the responses it has consumed
    auto next_item = [&](int item, uint32_t& cons, bool whole_warp) -> int {
        if (!dyn) {
            const int n = item + ncl;
            return n < n_items ? n : -1;
        }
        const int slot = (int)(cons % kClcDepth);
        const uint32_t ph = (cons / kClcDepth) & 1u;
        ptx::mbar_wait(&clc_full[slot], ph);
        const int x = ptx2::clc_decode(&clc_resp[slot]);
        if (whole_warp) __syncwarp();
        if (!whole_warp || lane == 0) ptx2::mbar_arrive_cluster(&clc_empty[slot], 0);
        ++cons;
        return x < 0 ? -1 : (x >> 1);
    };

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < p.nstages; ++s) {
            ptx::mbar_init(&full[s], 2);   // leader's arrive.expect_tx + the peer's remote arrive
            ptx::mbar_init(&empty[s], 1);  // multicast tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&tmem_full[s], 1);   // multicast tcgen05.commit
            ptx::mbar_init(&tmem_empty[s], 2 * kGemm2EpiWarps);  // 8 epilogue warps x 2 CTAs
        }
        for (int s = 0; s < kClcDepth; ++s) {
            ptx::mbar_init(&clc_full[s], 1);  // the scheduler's arrive.expect_tx (+ 16 response bytes)
            // leader: TMA lane, MMA warp, 8 epilogue warps; peer: TMA lane, 8 epilogue warps
            ptx::mbar_init(&clc_empty[s], 2 * kGemm2EpiWarps + 3);
        }
        ptx::fence_barrier_init();
    }
    ptx2::cluster_sync();  // barriers of both CTAs initialised before any remote arrive / 2-SM alloc
    if (warp == 1) {
        ptx2::tmem_alloc_2sm(tmem_base_smem, kTmemCols);
        ptx2::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            const uint64_t pol_a = (p.n_groups > 1) ? ptx::l2_policy_evict_last() : ptx::l2_policy_evict_first();
            const uint64_t pol_b = ptx::l2_policy_evict_last();
            const uint32_t pair_bytes = 2u * (uint32_t)(kABytes + b_bytes);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t cons = 0, fetch = 0;
            for (int item = cid; item >= 0; item = next_item(item, cons, false)) {
                if (dyn && leader) {
                    // the scheduler: ask for the item after this one before loading this one's operands
                    const int slot = (int)(fetch % kClcDepth);
                    ptx::mbar_wait(&clc_empty[slot], ((fetch / kClcDepth) & 1u) ^ 1u);
                    ptx::mbar_arrive_expect_tx(&clc_full[slot], 16);
                    ptx2::mbar_arrive_expect_tx_cluster(&clc_full[slot], 1, 16);
                    ptx2::clc_try_cancel_multicast(&clc_resp[slot], &clc_full[slot]);
                    ++fetch;
                }
                const int tile = item / p.n_groups;
                const int group = item - tile * p.n_groups;
                const int row0 = (int)(scan_tile(p, tile) * (2 * kTileM)) + (int)rank * kTileM;
                const int q0 = group * p.nt + (int)rank * half_nt;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    ptx::mbar_wait(&empty[stage], phase ^ 1u);
                    uint8_t* sa = stages + (size_t)stage * p.stage_bytes;
                    uint8_t* sb = sa + kABytes;
                    if (leader) ptx::mbar_arrive_expect_tx(&full[stage], pair_bytes);
                    else ptx2::mbar_arrive_cluster(&full[stage], 0);
                    ptx2::tma_load_2d_2sm(sa, &tmap_a, kb * kBlockK, row0, &full[stage], pol_a);
                    ptx2::tma_load_2d_2sm(sb, &tmap_b, kb * kBlockK, q0, &full[stage], pol_b);
                    if (++stage == p.nstages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const uint32_t stage0_lo = (ptx::smem_u32(stages) >> 4);
            const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;
            constexpr uint32_t kDescHi = (uint32_t)(1024u >> 4) | (1u << 14) | (2u << 29);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            uint32_t cons = 0;
            for (int item = cid; item >= 0; item = next_item(item, cons, true), ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccStride);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    ptx::mbar_wait(&full[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_lo = (stage0_lo + (uint32_t)stage * stage_step) & 0x3fffu;
                    const uint32_t b_lo = (stage0_lo + (uint32_t)stage * stage_step + (kABytes >> 4)) & 0x3fffu;
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            const uint64_t da = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + 2u * k);
                            const uint64_t db = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + 2u * k);
                            ptx2::umma_bf16_2sm(d_tmem, da, db, p.idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        ptx2::umma_commit_2sm(&empty[stage]);
                        if (kb == p.num_kb - 1) ptx2::umma_commit_2sm(&tmem_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == p.nstages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9 of both CTAs) =====================
        // Two warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31): warps 2-5 take the
        // first half of the query columns, warps 6-9 the second.  With one epilogue warp per scheduler the
        // epilogue ran at ~0.2 IPC (every dependent-issue latency exposed) and took as long as the MMAs of a
        // tile; two warps per scheduler hide each other's latencies.
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        uint2* stg = stage_buf + (size_t)(warp - 2) * kStageCap2;
        int it = 0;
        uint32_t cons = 0;
        for (int item = cid; item >= 0; item = next_item(item, cons, true), ++it) {
            const int tile = item / p.n_groups;
            const int group = item - tile * p.n_groups;
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            const int64_t row_warp0 = scan_tile(p, tile) * (2 * kTileM) + (int64_t)rank * kTileM + quarter * 32;
            const int64_t dense_slot0 = (int64_t)tile * (2 * kTileM) + (int64_t)rank * kTileM + quarter * 32;
            const int q0 = group * p.nt + half * half_nt;
            int ncols = p.batch - q0;
            if (ncols > half_nt) ncols = half_nt;
            if (ncols < 0) ncols = 0;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                   (uint32_t)(acc * kAccStride + half * half_nt);
            uint64_t* rel = &tmem_empty[acc];
            epilogue_item(p, taddr, row_warp0, dense_slot0, lane, q0, ncols, half_nt, stg, kStageCap2,
                          [rel, lane]() { if (lane == 0) ptx2::mbar_arrive_cluster(rel, 0); });
        }
    }

    // no CTA may exit (or free its TMEM) while its peer can still signal its barriers / read its smem
    __syncwarp();
    ptx::tc_fence_before();
    ptx2::cluster_sync();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx2::tmem_dealloc_2sm(tmem_base, kTmemCols);
    }
}

int encode_2d(CUtensorMap* out, const void* base, int64_t rows, int dim, int box_rows, bool half_tiles);  // gemm.cu
void set_scan_order(GemmParams& p, const GemmArgs& a, int tile_rows);                     // gemm.cu

int launch_gemm_2cta(const GemmArgs& a, cudaStream_t stream) {
    const Store* s = a.store;
    GemmParams p;
    p.dim = s->dim;
    p.tf32 = 0;
    p.kb_elems = kBlockK;
    p.num_kb = (s->dim + kBlockK - 1) / kBlockK;
    p.nt = gemm_group_width(a.bpad);
    CMW_REQUIRE(p.nt % 64 == 0 && a.bpad % p.nt == 0, "launch_gemm_2cta: bad query padding %d", a.bpad);
    p.n_groups = a.bpad / p.nt;
    p.batch = a.batch;
    p.row_begin = a.row_begin;
    p.row_end = a.row_end;
    p.n_tiles = (int)((a.row_end - a.row_begin + 2 * kTileM - 1) / (2 * kTileM));
    set_scan_order(p, a, 2 * kTileM);
    p.stage_bytes = kABytes + (p.nt / 2) * kBlockK * 2;
    const size_t tail = (2 * kMaxStages + 4) * sizeof(uint64_t) + 64 + kGemm2EpiWarps * kStageCap2 * sizeof(uint2) +
                        kClcDepth * (sizeof(uint4) + 2 * sizeof(uint64_t));
    int nst = (int)((220 * 1024 - tail - 1024) / (size_t)p.stage_bytes);
    if (nst > kMaxStages) nst = kMaxStages;
    p.nstages = nst;
    p.dense = a.dense;
    p.dense_scores = a.wide_scores ? a.wide_scores : a.pool.scores;
    p.dense_ids = a.wide_scores ? a.wide_ids : a.pool.ids;
    p.dense_stride = a.wide_scores ? a.wide_stride : kPoolCap;
    // instruction descriptor: D = f32, A / B format (0 = fp16, 1 = bf16), K-major, N >> 3 at bit 17, M = 256 >> 4 at bit 24
    const uint32_t fmt = s->half_tiles ? 0u : 1u;
    p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.nt >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    p.row_mul = a.row_mul;
    p.pool_scores = a.pool.scores;
    p.pool_ids = a.pool.ids;
    p.pool_cnt = a.pool.cnt;
    p.pool_thr = a.pool.thr;
    CUtensorMap tmap_b;
    int rc = encode_2d(&tmap_b, a.q_bf16, a.bpad, s->dim, p.nt / 2, s->half_tiles);
    if (rc) return rc;
    const size_t smem = (size_t)nst * p.stage_bytes + tail + 1024;
    static SmemAttrCache smem_set;
    if (smem_set.needs(smem)) {
        CMW_CUDA_OK(cudaFuncSetAttribute(gemm_topk_kernel_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        smem_set.done(smem);
    }
    const int n_items = p.n_tiles * p.n_groups;
    int clusters = s->sm_count / 2;
    if (n_items < clusters) clusters = n_items;
    // dynamic scheduling (cluster launch control): one pair per item in the grid, the resident pairs steal the rest
    p.dynamic = (a.opt->gemm_clc != 0 && n_items > clusters) ? 1 : 0;
    if (p.dynamic) clusters = n_items;
    gemm_topk_kernel_2cta<<<2 * clusters, kGemm2Threads, smem, stream>>>(s->tmap_bf16, tmap_b, p);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace cmw
