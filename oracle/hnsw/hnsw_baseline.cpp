// hnsw_baseline.cpp -- hnswlib-equivalent HNSW index: the CPU baseline of the dense-retrieval path.
// TEST / BENCHMARK INFRASTRUCTURE ONLY -- nothing under cmw_rag_b200/ may link or call this.
//
// Why it exists: the reference (arterm-sedov/cmw-rag) answers similarity_search_async
// (rag_engine/storage/vector_store.py:54-66) through chromadb==1.3.0
// (rag_engine/requirements_frozen.txt:15), whose vector segment is an hnswlib index created with
// {"hnsw:space": "cosine"} (vector_store.py:48-51) and otherwise Chroma's defaults.  Neither
// chromadb nor hnswlib is installable offline, so the published algorithm (Malkov & Yashunin,
// "Efficient and robust approximate nearest neighbor search using Hierarchical Navigable Small
// World graphs", as implemented by hnswlib 0.7/0.8) is restated here with the defaults Chroma 1.x
// uses: M = 16 (max neighbours; 2M on layer 0), ef_construction = 100, ef_search = 100 (the layer-0
// beam is max(ef_search, k)), fp32 vectors, cosine = L2-normalise on insert/query then
// distance = 1 - <a, b>, level = floor(-ln(U) / ln(M)), neighbour selection by the "heuristic"
// rule (keep a candidate only if it is closer to the base point than to every neighbour already
// kept), bidirectional links with re-selection when a list overflows, multi-threaded inserts with
// per-node locks, one query per thread.  Every report labels it "hnswlib-equivalent
// re-implementation, chromadb 1.3.0 defaults assumed (not verifiable offline)".
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <queue>
#include <random>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct SpinLock {
    std::atomic_flag f = ATOMIC_FLAG_INIT;
    void lock() {
        while (f.test_and_set(std::memory_order_acquire)) {
        }
    }
    void unlock() { f.clear(std::memory_order_release); }
};

typedef std::pair<float, int32_t> Cand;  // (distance, id)

struct Index {
    int dim, M, M0, efc;
    int64_t cap;
    std::atomic<int64_t> count{0};
    double mult;
    float* vec = nullptr;             // [cap, dim] normalised
    int32_t* link0 = nullptr;         // [cap, M0 + 1]   (count, neighbours...)
    std::vector<int32_t*> upper;      // per element: [level][M + 1] or nullptr
    std::vector<int> level;
    std::vector<SpinLock> locks;
    std::mutex global;
    int32_t entry = -1;
    int maxlevel = -1;
    std::mt19937_64 rng;

    inline const float* v(int32_t i) const { return vec + (size_t)i * dim; }
    inline int32_t* l0(int32_t i) const { return link0 + (size_t)i * (M0 + 1); }
    inline int32_t* lu(int32_t i, int lev) const { return upper[i] + (size_t)(lev - 1) * (M + 1); }
    inline int32_t* links(int32_t i, int lev) const { return lev == 0 ? l0(i) : lu(i, lev); }
};

inline float dist_ip(const float* a, const float* b, int d) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = 0;
    const int d16 = d & ~63;
    for (; i < d16; i += 64) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma omp simd reduction(+ : t0)
        for (int j = 0; j < 16; ++j) t0 += a[i + j] * b[i + j];
#pragma omp simd reduction(+ : t1)
        for (int j = 16; j < 32; ++j) t1 += a[i + j] * b[i + j];
#pragma omp simd reduction(+ : t2)
        for (int j = 32; j < 48; ++j) t2 += a[i + j] * b[i + j];
#pragma omp simd reduction(+ : t3)
        for (int j = 48; j < 64; ++j) t3 += a[i + j] * b[i + j];
        s0 += t0;
        s1 += t1;
        s2 += t2;
        s3 += t3;
    }
    float s = 0.f;
#pragma omp simd reduction(+ : s)
    for (int j = i; j < d; ++j) s += a[j] * b[j];
    return 1.0f - (s + (s0 + s1) + (s2 + s3));
}

struct Visited {
    std::vector<uint16_t> tag;
    uint16_t cur = 0;
    void reset(int64_t n) {
        if ((int64_t)tag.size() < n) tag.assign(n, 0);
        if (++cur == 0) {
            std::fill(tag.begin(), tag.end(), 0);
            cur = 1;
        }
    }
};

// beam search on one layer; returns a max-heap (worst on top) of at most ef results
void search_layer(const Index& ix, const float* q, int32_t ep, float ep_dist, int lev, int ef, Visited& vis,
                  std::priority_queue<Cand>& top, bool lock_nodes) {
    std::priority_queue<Cand, std::vector<Cand>, std::greater<Cand>> cand;  // closest first
    top = std::priority_queue<Cand>();
    top.emplace(ep_dist, ep);
    cand.emplace(ep_dist, ep);
    vis.tag[ep] = vis.cur;
    float lower = ep_dist;
    std::vector<int32_t> nb;
    while (!cand.empty()) {
        Cand c = cand.top();
        if (c.first > lower && (int)top.size() >= ef) break;
        cand.pop();
        {
            Index& mix = const_cast<Index&>(ix);
            if (lock_nodes) mix.locks[c.second].lock();
            const int32_t* l = ix.links(c.second, lev);
            nb.assign(l + 1, l + 1 + l[0]);
            if (lock_nodes) mix.locks[c.second].unlock();
        }
        for (size_t ni = 0; ni < nb.size(); ++ni) {
            const int32_t n = nb[ni];
            // like hnswlib's searchBaseLayer: start fetching the next neighbour's vector while this one is scored
            if (ni + 1 < nb.size()) {
                const char* nx = reinterpret_cast<const char*>(ix.v(nb[ni + 1]));
                __builtin_prefetch(nx);
                __builtin_prefetch(nx + 64);
            }
            if (vis.tag[n] == vis.cur) continue;
            vis.tag[n] = vis.cur;
            const float d = dist_ip(q, ix.v(n), ix.dim);
            if ((int)top.size() < ef || d < lower) {
                cand.emplace(d, n);
                top.emplace(d, n);
                if ((int)top.size() > ef) top.pop();
                lower = top.top().first;
            }
        }
    }
}

// hnswlib getNeighborsByHeuristic2: candidates closest-first; keep c if dist(c, base) < dist(c, kept) for all kept
void select_heuristic(const Index& ix, std::vector<Cand>& sorted_cands, int m, std::vector<int32_t>& out) {
    out.clear();
    if ((int)sorted_cands.size() <= m) {
        for (auto& c : sorted_cands) out.push_back(c.second);
        return;
    }
    for (auto& c : sorted_cands) {
        if ((int)out.size() >= m) break;
        bool good = true;
        for (int32_t s : out) {
            if (dist_ip(ix.v(s), ix.v(c.second), ix.dim) < c.first) {
                good = false;
                break;
            }
        }
        if (good) out.push_back(c.second);
    }
}

void connect(Index& ix, int32_t cur, std::priority_queue<Cand>& top, int lev) {
    const int mmax = lev == 0 ? ix.M0 : ix.M;
    std::vector<Cand> sorted;
    while (!top.empty()) {
        sorted.push_back(top.top());
        top.pop();
    }
    std::reverse(sorted.begin(), sorted.end());  // closest first
    std::vector<int32_t> sel;
    select_heuristic(ix, sorted, ix.M, sel);
    {
        ix.locks[cur].lock();
        int32_t* l = ix.links(cur, lev);
        l[0] = (int32_t)sel.size();
        for (size_t i = 0; i < sel.size(); ++i) l[1 + i] = sel[i];
        ix.locks[cur].unlock();
    }
    std::vector<Cand> cands;
    std::vector<int32_t> resel;
    for (int32_t nb : sel) {
        ix.locks[nb].lock();
        int32_t* l = ix.links(nb, lev);
        bool present = false;
        for (int i = 0; i < l[0]; ++i) present |= (l[1 + i] == cur);
        if (!present) {
            if (l[0] < mmax) {
                l[1 + l[0]] = cur;
                l[0] += 1;
            } else {
                const float dmax = dist_ip(ix.v(cur), ix.v(nb), ix.dim);
                cands.clear();
                cands.emplace_back(dmax, cur);
                for (int i = 0; i < l[0]; ++i) cands.emplace_back(dist_ip(ix.v(l[1 + i]), ix.v(nb), ix.dim), l[1 + i]);
                std::sort(cands.begin(), cands.end());
                select_heuristic(ix, cands, mmax, resel);
                l[0] = (int32_t)resel.size();
                for (size_t i = 0; i < resel.size(); ++i) l[1 + i] = resel[i];
            }
        }
        ix.locks[nb].unlock();
    }
}

void add_point(Index& ix, int32_t cur, Visited& vis) {
    int lev;
    {
        std::lock_guard<std::mutex> g(ix.global);
        std::uniform_real_distribution<double> u(0.0, 1.0);
        double r = -log(std::max(u(ix.rng), 1e-300)) * ix.mult;
        lev = (int)r;
        ix.level[cur] = lev;
        if (lev > 0) {
            ix.upper[cur] = (int32_t*)calloc((size_t)lev * (ix.M + 1), sizeof(int32_t));
        }
    }
    std::unique_lock<std::mutex> glock(ix.global);
    const int maxl = ix.maxlevel;
    int32_t ep = ix.entry;
    if (lev <= maxl) glock.unlock();
    if (ep < 0) {
        ix.entry = cur;
        ix.maxlevel = lev;
        return;
    }
    const float* q = ix.v(cur);
    float d = dist_ip(q, ix.v(ep), ix.dim);
    for (int l = maxl; l > lev; --l) {
        bool changed = true;
        while (changed) {
            changed = false;
            ix.locks[ep].lock();
            const int32_t* ln = ix.links(ep, l);
            std::vector<int32_t> nb(ln + 1, ln + 1 + ln[0]);
            ix.locks[ep].unlock();
            for (int32_t n : nb) {
                const float dn = dist_ip(q, ix.v(n), ix.dim);
                if (dn < d) {
                    d = dn;
                    ep = n;
                    changed = true;
                }
            }
        }
    }
    std::priority_queue<Cand> top;
    for (int l = std::min(lev, maxl); l >= 0; --l) {
        vis.reset(ix.cap);
        search_layer(ix, q, ep, d, l, ix.efc, vis, top, true);
        // next layer's entry point: the closest found
        std::priority_queue<Cand> copy = top;
        Cand best = copy.top();
        while (!copy.empty()) {
            if (copy.top().first < best.first) best = copy.top();
            copy.pop();
        }
        connect(ix, cur, top, l);
        ep = best.second;
        d = best.first;
    }
    if (lev > maxl) {
        ix.entry = cur;
        ix.maxlevel = lev;
    }
}

}  // namespace

extern "C" {

int hnsw_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void* hnsw_create(int dim, int64_t max_elems, int M, int ef_construction, uint64_t seed) {
    Index* ix = new Index();
    ix->dim = dim;
    ix->M = M;
    ix->M0 = 2 * M;
    ix->efc = ef_construction;
    ix->cap = max_elems;
    ix->mult = 1.0 / log((double)M);
    ix->vec = (float*)malloc((size_t)max_elems * dim * sizeof(float));
    ix->link0 = (int32_t*)calloc((size_t)max_elems * (ix->M0 + 1), sizeof(int32_t));
    ix->upper.assign(max_elems, nullptr);
    ix->level.assign(max_elems, 0);
    ix->locks = std::vector<SpinLock>(max_elems);
    ix->rng.seed(seed);
    if (!ix->vec || !ix->link0) {
        delete ix;
        return nullptr;
    }
    return ix;
}

void hnsw_free(void* h) {
    Index* ix = (Index*)h;
    if (!ix) return;
    for (auto p : ix->upper) free(p);
    free(ix->vec);
    free(ix->link0);
    delete ix;
}

// rows: [n, dim]; normalised here (cosine space) -- returns the number of elements in the index
int64_t hnsw_add(void* h, const float* rows, int64_t n, int nthreads) {
    Index& ix = *(Index*)h;
    const int64_t base = ix.count.load();
    if (base + n > ix.cap) return -1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* src = rows + (size_t)i * ix.dim;
        float* dst = ix.vec + (size_t)(base + i) * ix.dim;
        double s = 0.0;
        for (int d = 0; d < ix.dim; ++d) s += (double)src[d] * src[d];
        const float inv = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.f;
        for (int d = 0; d < ix.dim; ++d) dst[d] = src[d] * inv;
    }
    // the first point is inserted alone (it becomes the entry point)
    int64_t start = 0;
    if (base == 0 && n > 0) {
        Visited vis;
        vis.reset(ix.cap);
        add_point(ix, 0, vis);
        start = 1;
    }
#pragma omp parallel num_threads(nthreads)
    {
        Visited vis;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = start; i < n; ++i) add_point(ix, (int32_t)(base + i), vis);
    }
    ix.count.store(base + n);
    return base + n;
}

// queries: [nq, dim] (normalised here); ids [nq, k] (-1 padded), dists [nq, k] = 1 - cosine
void hnsw_search(void* h, const float* queries, int nq, int k, int ef_search, int64_t* ids, float* dists,
                 int nthreads) {
    Index& ix = *(Index*)h;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    const int ef = std::max(ef_search, k);
#pragma omp parallel num_threads(nthreads)
    {
        Visited vis;
        std::vector<float> qn(ix.dim);
        std::priority_queue<Cand> top;
#pragma omp for schedule(dynamic, 1)
        for (int qi = 0; qi < nq; ++qi) {
            const float* src = queries + (size_t)qi * ix.dim;
            double s = 0.0;
            for (int d = 0; d < ix.dim; ++d) s += (double)src[d] * src[d];
            const float inv = s > 0.0 ? (float)(1.0 / sqrt(s)) : 0.f;
            for (int d = 0; d < ix.dim; ++d) qn[d] = src[d] * inv;
            for (int j = 0; j < k; ++j) {
                ids[(size_t)qi * k + j] = -1;
                dists[(size_t)qi * k + j] = INFINITY;
            }
            if (ix.entry < 0) continue;
            int32_t ep = ix.entry;
            float d = dist_ip(qn.data(), ix.v(ep), ix.dim);
            for (int l = ix.maxlevel; l > 0; --l) {
                bool changed = true;
                while (changed) {
                    changed = false;
                    const int32_t* ln = ix.links(ep, l);
                    for (int i = 0; i < ln[0]; ++i) {
                        const int32_t n = ln[1 + i];
                        const float dn = dist_ip(qn.data(), ix.v(n), ix.dim);
                        if (dn < d) {
                            d = dn;
                            ep = n;
                            changed = true;
                        }
                    }
                }
            }
            vis.reset(ix.cap);
            search_layer(ix, qn.data(), ep, d, 0, ef, vis, top, false);
            while ((int)top.size() > k) top.pop();
            int j = (int)top.size() - 1;
            while (!top.empty()) {
                ids[(size_t)qi * k + j] = top.top().second;
                dists[(size_t)qi * k + j] = top.top().first;
                top.pop();
                --j;
            }
        }
    }
}

}  // extern "C"
