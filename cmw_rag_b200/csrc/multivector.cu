// multivector.cu -- K4: segmented union / first-seen dedup / pre-rerank cap / kbId group reduction.
//
// Replaces, for Q long queries at once, the Python loops of the reference's
// RAGRetriever.retrieve_async (rag_engine/retrieval/retriever.py):
//   :185-194  ordered union over segments (segment-major, rank-minor), first occurrence wins,
//             key = metadata["stable_id"] (one per corpus row -> the row id here)
//   :208-210  candidates[:pre_rerank_limit]
//   :229-231  scored_candidates[:top_k]                       (`limit`)
//   :234-242  group by extract_numeric_kbid(kbId) or str(kbId) (precomputed dense kb_gid per row,
//             rag_engine/utils/metadata_utils.py:20-32), skip falsy kbId, keep max score,
//             dict order = first appearance
//   :307      articles.sort(key=score, reverse=True)  -- stable
//
// One CTA per long query; everything is integer/compare work in shared memory (bitonic sorts +
// block scans), checked bit for bit against the CPU restatement in the parity tests.
#include "common.cuh"

namespace cmw {

constexpr int kMvMaxEntries = 2048;
constexpr int kMvThreads = 512;

// exclusive scan of data[0..n) in place; returns the total.  All threads must call.
__device__ int block_excl_scan(int* data, int n, int* warp_sums) {
    __syncthreads();
    const int T = blockDim.x;
    const int per = (n + T - 1) / T;
    const int begin = threadIdx.x * per;
    const int end = (begin + per < n) ? begin + per : n;
    int sum = 0;
    for (int i = begin; i < end; ++i) sum += data[i];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = T >> 5;
    int v = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warp_sums[w] = v;
    __syncthreads();
    if (w == 0) {
        int ws = lane < nw ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += t;
        }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    int run = (w > 0 ? warp_sums[w - 1] : 0) + v - sum;
    for (int i = begin; i < end; ++i) {
        int t = data[i];
        data[i] = run;
        run += t;
    }
    const int total = warp_sums[nw - 1];
    __syncthreads();
    return total;
}

struct MvParams {
    const int32_t* kb_gid;
    int64_t kb_rows;
    int64_t id_offset;
    const int64_t* ids;
    const float* scores;
    int S, k, P, limit;
    int64_t* cand_ids;
    float* cand_scores;
    float* cand_best;
    int32_t* cand_n;
    int32_t* cand_grp;
    int32_t* grp_gid;
    float* grp_max;
    int32_t* grp_cnt;
    int32_t* grp_first;
    int32_t* grp_order;
    int32_t* grp_n;
};

__global__ void __launch_bounds__(kMvThreads) multivector_kernel(const MvParams p) {
    extern __shared__ __align__(16) uint8_t mv_smem[];
    const int q = blockIdx.x;
    const int n = p.S * p.k;
    const int n2 = next_pow2(n < 2 ? 2 : n);
    const int P = p.P;
    // shared-memory carve-up (n2 entries each)
    uint64_t* hi = reinterpret_cast<uint64_t*>(mv_smem);   // sort keys (ids, later group keys)
    uint64_t* lo = hi + n2;
    int64_t* c_id = reinterpret_cast<int64_t*>(lo + n2);   // candidate ids by candidate index
    float* sc_pos = reinterpret_cast<float*>(c_id + n2);   // scores by position
    float* best_at = sc_pos + n2;                          // max over occurrences, by first position
    float* c_sc = best_at + n2;                            // first-seen score by candidate index
    float* g_max_at = c_sc + n2;                           // by first candidate index of the group
    float* g_max = g_max_at + n2;                          // by group index
    int* flag = reinterpret_cast<int*>(g_max + n2);        // first-occurrence flags -> candidate index
    int* g_flag = flag + n2;                               // group-head flags -> group index
    int* g_cnt_at = g_flag + n2;
    int* g_gid_at = g_cnt_at + n2;
    int* g_first_of = g_gid_at + n2;                       // candidate -> first candidate of its group
    int* warp_sums = g_first_of + n2;                      // [32]

    const int64_t* ids = p.ids + (size_t)q * n;
    const float* scores = p.scores + (size_t)q * n;

    // ---- A: sort (id, position) -----------------------------------------------------------
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        uint64_t h = ~0ull, l = ~0ull;
        float s = -INFINITY;
        if (i < n) {
            const int64_t id = ids[i];
            s = scores[i];
            if (id >= 0) {
                h = (uint64_t)id;
                l = (uint64_t)i;
            }
        }
        hi[i] = h;
        lo[i] = l;
        sc_pos[i] = s;
        flag[i] = 0;
        best_at[i] = -INFINITY;
    }
    bitonic_sort_u128(hi, lo, n2);

    // ---- B: run heads = first occurrences; best = max over the run ---------------------------
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const uint64_t h = hi[i];
        if (h == ~0ull && lo[i] == ~0ull) continue;
        if (i > 0 && hi[i - 1] == h) continue;
        const int first_pos = (int)lo[i];
        float best = sc_pos[first_pos];
        for (int j = i + 1; j < n2 && hi[j] == h; ++j) {
            const float s = sc_pos[(int)lo[j]];
            if (s > best) best = s;
        }
        flag[first_pos] = 1;
        best_at[first_pos] = best;
    }
    __syncthreads();
    // ---- C: candidate index = rank of the first occurrence in position order -------------------
    // (flag becomes the exclusive scan; remember which positions were heads through best_at/ids)
    for (int i = threadIdx.x; i < n2; i += blockDim.x) g_flag[i] = flag[i];  // keep a copy of the flags
    const int uniq = block_excl_scan(flag, n2, warp_sums);
    const int cand_n = uniq < P ? uniq : P;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (g_flag[i]) {
            const int ci = flag[i];
            if (ci < P) {
                c_id[ci] = ids[i];
                c_sc[ci] = sc_pos[i];
                if (p.cand_ids) p.cand_ids[(size_t)q * P + ci] = ids[i];
                if (p.cand_scores) p.cand_scores[(size_t)q * P + ci] = sc_pos[i];
                if (p.cand_best) p.cand_best[(size_t)q * P + ci] = best_at[i];
            }
        }
    }
    for (int ci = cand_n + threadIdx.x; ci < P; ci += blockDim.x) {
        if (p.cand_ids) p.cand_ids[(size_t)q * P + ci] = -1;
        if (p.cand_scores) p.cand_scores[(size_t)q * P + ci] = -INFINITY;
        if (p.cand_best) p.cand_best[(size_t)q * P + ci] = -INFINITY;
    }
    if (threadIdx.x == 0 && p.cand_n) p.cand_n[q] = cand_n;
    __syncthreads();

    // ---- D: group by kb_gid over the first L candidates -----------------------------------------
    const int L = (p.limit > 0 && p.limit < cand_n) ? p.limit : cand_n;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < L) {
            const int64_t local = c_id[i] - p.id_offset;
            int32_t g = -1;
            if (local >= 0 && local < p.kb_rows) g = p.kb_gid[local];
            if (g >= 0) key = ((uint64_t)(uint32_t)g << 32) | (uint64_t)i;
        }
        hi[i] = key;
        g_flag[i] = 0;
        g_first_of[i] = -1;
    }
    bitonic_sort_u64(hi, n2);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const uint64_t key = hi[i];
        if (key == ~0ull) continue;
        const uint32_t g = (uint32_t)(key >> 32);
        if (i > 0 && (uint32_t)(hi[i - 1] >> 32) == g) continue;
        const int first_ci = (int)(uint32_t)key;
        float mx = -INFINITY;
        int cnt = 0;
        for (int j = i; j < n2 && hi[j] != ~0ull && (uint32_t)(hi[j] >> 32) == g; ++j) {
            const int ci = (int)(uint32_t)hi[j];
            const float s = c_sc[ci];
            if (s > mx) mx = s;
            ++cnt;
            g_first_of[ci] = first_ci;
        }
        g_flag[first_ci] = 1;
        g_max_at[first_ci] = mx;
        g_cnt_at[first_ci] = cnt;
        g_gid_at[first_ci] = (int)g;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += blockDim.x) flag[i] = g_flag[i];
    const int grp_n = block_excl_scan(flag, n2, warp_sums);  // flag[ci] = group index of a head at ci
    for (int ci = threadIdx.x; ci < P; ci += blockDim.x) {
        int gi = -1;
        if (ci < L && g_first_of[ci] >= 0) gi = flag[g_first_of[ci]];
        if (p.cand_grp) p.cand_grp[(size_t)q * P + ci] = gi;
        if (ci < L && g_flag[ci]) {
            const int g = flag[ci];
            g_max[g] = g_max_at[ci];
            if (p.grp_gid) p.grp_gid[(size_t)q * P + g] = g_gid_at[ci];
            if (p.grp_max) p.grp_max[(size_t)q * P + g] = g_max_at[ci];
            if (p.grp_cnt) p.grp_cnt[(size_t)q * P + g] = g_cnt_at[ci];
            if (p.grp_first) p.grp_first[(size_t)q * P + g] = ci;
        }
    }
    for (int g = grp_n + threadIdx.x; g < P; g += blockDim.x) {
        if (p.grp_gid) p.grp_gid[(size_t)q * P + g] = -1;
        if (p.grp_max) p.grp_max[(size_t)q * P + g] = -INFINITY;
        if (p.grp_cnt) p.grp_cnt[(size_t)q * P + g] = 0;
        if (p.grp_first) p.grp_first[(size_t)q * P + g] = -1;
    }
    if (threadIdx.x == 0 && p.grp_n) p.grp_n[q] = grp_n;
    __syncthreads();

    // ---- E: stable score-descending order of the groups ------------------------------------------
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < grp_n) {
            float s = g_max[i];
            if (s == 0.f) s = 0.f;  // -0.0 and +0.0 compare equal in the reference's sort
            key = desc_key(s, (uint32_t)i);
        }
        hi[i] = key;
    }
    bitonic_sort_u64(hi, n2);
    if (p.grp_order) {
        for (int g = threadIdx.x; g < P; g += blockDim.x)
            p.grp_order[(size_t)q * P + g] = (g < grp_n) ? (int)(uint32_t)hi[g] : -1;
    }
}

}  // namespace cmw

using namespace cmw;

extern "C" int cmw_multivector(const int32_t* kb_gid_dev, int64_t kb_rows, int64_t id_offset,
                               const int64_t* ids_dev, const float* scores_dev, int Q, int S, int k,
                               int prl, int limit, int64_t* cand_ids, float* cand_scores,
                               float* cand_best, int32_t* cand_n, int32_t* cand_grp,
                               int32_t* grp_gid, float* grp_max, int32_t* grp_cnt,
                               int32_t* grp_first, int32_t* grp_order, int32_t* grp_n, void* stream) {
    CMW_REQUIRE(ids_dev && scores_dev, "cmw_multivector: ids/scores are NULL");
    CMW_REQUIRE(Q >= 0 && S >= 1 && k >= 1, "cmw_multivector: bad sizes");
    CMW_REQUIRE(kb_gid_dev != nullptr || kb_rows == 0, "cmw_multivector: kb_gid is NULL");
    if (Q == 0) return 0;
    const int n = S * k;
    CMW_REQUIRE(n <= kMvMaxEntries, "cmw_multivector: S*k = %d exceeds %d", n, kMvMaxEntries);
    MvParams p;
    p.kb_gid = kb_gid_dev;
    p.kb_rows = kb_rows;
    p.id_offset = id_offset;
    p.ids = ids_dev;
    p.scores = scores_dev;
    p.S = S;
    p.k = k;
    p.P = (prl > 0 && prl < n) ? prl : n;
    p.limit = limit;
    p.cand_ids = cand_ids;
    p.cand_scores = cand_scores;
    p.cand_best = cand_best;
    p.cand_n = cand_n;
    p.cand_grp = cand_grp;
    p.grp_gid = grp_gid;
    p.grp_max = grp_max;
    p.grp_cnt = grp_cnt;
    p.grp_first = grp_first;
    p.grp_order = grp_order;
    p.grp_n = grp_n;
    const int n2 = next_pow2_host(n < 2 ? 2 : n);
    const size_t smem = (size_t)n2 * (3 * 8 + 5 * 4 + 5 * 4) + 32 * sizeof(int) + 16;
    static SmemAttrCache smem_set;
    if (smem > 48 * 1024 && smem_set.needs(smem)) {
        CMW_CUDA_OK(cudaFuncSetAttribute(multivector_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set.done(smem);
    }
    multivector_kernel<<<Q, kMvThreads, smem, (cudaStream_t)stream>>>(p);
    CMW_LAUNCHED();
    CMW_CUDA_OK(cudaGetLastError());
    return 0;
}
