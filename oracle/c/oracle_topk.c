/* CPU oracle: exact cosine / inner-product top-k with fp64 accumulation.
 * TEST INFRASTRUCTURE ONLY -- nothing under cmw_rag_b200/ may link or call this.
 *
 * Restates the query of rag_engine/storage/vector_store.py:54-66 (k nearest rows in
 * the space declared at :48-51, best first) exactly, as BASELINE.json:north_star
 * prescribes; the arithmetic the reference delegates to chromadb==1.3.0 / hnswlib
 * (not in the tree, approximate) is replaced by its exact limit.
 *
 *   dot(q,c)  = fp64 sum of fp64(q[d])*fp64(c[d])  (products of two fp32 values are
 *               exact in fp64, so only the summation order matters; it is fixed below:
 *               8 interleaved partial sums, combined as a balanced tree)
 *   cosine    = dot(q,c) / sqrt(dot(q,q)*dot(c,c))   (0 when a norm is 0)
 *   order     = score descending, id ascending
 *
 * The summation order is independent of the row's position, so bit-identical rows get
 * bit-identical scores and ties fall to the lower id.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_COSINE 0
#define ORACLE_IP 1

typedef struct {
    double score;
    int64_t id;
} cand_t;

/* a is worse than b  <=>  a sorts after b in (score desc, id asc) */
static inline int worse(const cand_t* a, const cand_t* b) {
    if (a->score != b->score) return a->score < b->score;
    return a->id > b->id;
}

/* heap with the WORST element at the root */
static void heap_sift_down(cand_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && worse(&h[l], &h[w])) w = l;
        if (r < n && worse(&h[r], &h[w])) w = r;
        if (w == i) return;
        cand_t t = h[i];
        h[i] = h[w];
        h[w] = t;
        i = w;
    }
}
static void heap_sift_up(cand_t* h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!worse(&h[i], &h[p])) return;
        cand_t t = h[i];
        h[i] = h[p];
        h[p] = t;
        i = p;
    }
}
static inline void heap_offer(cand_t* h, int* n, int k, cand_t c) {
    if (*n < k) {
        h[*n] = c;
        heap_sift_up(h, *n);
        (*n)++;
    } else if (worse(&h[0], &c)) {
        h[0] = c;
        heap_sift_down(h, k, 0);
    }
}

static int cmp_best_first(const void* pa, const void* pb) {
    const cand_t* a = (const cand_t*)pa;
    const cand_t* b = (const cand_t*)pb;
    if (worse(a, b)) return 1;
    if (worse(b, a)) return -1;
    return 0;
}

__attribute__((target_clones("avx512f", "avx2", "default")))
static double dot64(const double* q, const double* c, int d) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    for (; i + 8 <= d; i += 8)
        for (int j = 0; j < 8; ++j) acc[j] += q[i + j] * c[i + j];
    for (int j = 0; i < d; ++i, ++j) acc[j] += q[i] * c[i];
    return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

double oracle_dot64_f32(const float* q, const float* c, int d) {
    double* qq = (double*)malloc(sizeof(double) * d * 2);
    double* cc = qq + d;
    for (int i = 0; i < d; ++i) {
        qq[i] = q[i];
        cc[i] = c[i];
    }
    double r = dot64(qq, cc, d);
    free(qq);
    return r;
}

/* corpus f32[n,d], queries f32[b,d], live u8[n] or NULL.
 * out_ids i64[b,k], out_scores f64[b,k]; unused slots: id -1, score -inf.  Returns 0. */
int oracle_topk_f32(const float* corpus, int64_t n, int d, const float* queries, int b, int k,
                    int metric, const uint8_t* live, int64_t id_offset, int64_t* out_ids,
                    double* out_scores, int nthreads) {
    if (n < 0 || d <= 0 || b < 0 || k <= 0) return -1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    double* q64 = (double*)malloc(sizeof(double) * (size_t)b * d);
    double* qn = (double*)malloc(sizeof(double) * (size_t)(b > 0 ? b : 1));
    for (int i = 0; i < b; ++i) {
        for (int j = 0; j < d; ++j) q64[(size_t)i * d + j] = queries[(size_t)i * d + j];
        qn[i] = dot64(q64 + (size_t)i * d, q64 + (size_t)i * d, d);
    }
    cand_t* heaps = (cand_t*)malloc(sizeof(cand_t) * (size_t)nthreads * b * k);
    int* hn = (int*)calloc((size_t)nthreads * (b > 0 ? b : 1), sizeof(int));

#pragma omp parallel num_threads(nthreads)
    {
#ifdef _OPENMP
        int t = omp_get_thread_num();
#else
        int t = 0;
#endif
        double* c64 = (double*)malloc(sizeof(double) * d);
        cand_t* myheap = heaps + (size_t)t * b * k;
        int* myn = hn + (size_t)t * b;
#pragma omp for schedule(static)
        for (int64_t r = 0; r < n; ++r) {
            if (live && !live[r]) continue;
            const float* c = corpus + (size_t)r * d;
            for (int j = 0; j < d; ++j) c64[j] = c[j];
            double cn = (metric == ORACLE_COSINE) ? dot64(c64, c64, d) : 1.0;
            for (int i = 0; i < b; ++i) {
                double s = dot64(q64 + (size_t)i * d, c64, d);
                if (metric == ORACLE_COSINE) {
                    double den = sqrt(qn[i] * cn);
                    s = den > 0 ? s / den : 0.0;
                }
                cand_t cd = {s, r + id_offset};
                heap_offer(myheap + (size_t)i * k, &myn[i], k, cd);
            }
        }
        free(c64);
    }

    cand_t* all = (cand_t*)malloc(sizeof(cand_t) * (size_t)nthreads * k);
    for (int i = 0; i < b; ++i) {
        int m = 0;
        for (int t = 0; t < nthreads; ++t) {
            int c = hn[(size_t)t * b + i];
            memcpy(all + m, heaps + ((size_t)t * b + i) * k, sizeof(cand_t) * c);
            m += c;
        }
        qsort(all, m, sizeof(cand_t), cmp_best_first);
        for (int j = 0; j < k; ++j) {
            if (j < m) {
                out_ids[(size_t)i * k + j] = all[j].id;
                out_scores[(size_t)i * k + j] = all[j].score;
            } else {
                out_ids[(size_t)i * k + j] = -1;
                out_scores[(size_t)i * k + j] = -INFINITY;
            }
        }
    }
    free(all);
    free(hn);
    free(heaps);
    free(qn);
    free(q64);
    return 0;
}

/* Exact top-k restricted to a candidate list per query: cand i64[b,m] (negative = skip).  The same fp64
 * arithmetic and order as oracle_topk_f32, so on candidate lists that contain the true top-k (a prefilter with a
 * verified margin, see bench.py) the result IS oracle_topk_f32's.  out_ids i64[b,k], out_scores f64[b,k]. */
int oracle_topk_candidates_f32(const float* corpus, int64_t n, int d, const float* queries, int b, int k,
                               int metric, const int64_t* cand, int m, int64_t* out_ids, double* out_scores,
                               int nthreads) {
    if (n < 0 || d <= 0 || b < 0 || k <= 0 || m <= 0) return -1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        double* q64 = (double*)malloc(sizeof(double) * d * 2);
        double* c64 = q64 + d;
        cand_t* all = (cand_t*)malloc(sizeof(cand_t) * (size_t)m);
#pragma omp for schedule(dynamic, 4)
        for (int i = 0; i < b; ++i) {
            for (int j = 0; j < d; ++j) q64[j] = queries[(size_t)i * d + j];
            const double qn = dot64(q64, q64, d);
            int cnt = 0;
            for (int e = 0; e < m; ++e) {
                const int64_t r = cand[(size_t)i * m + e];
                if (r < 0 || r >= n) continue;
                const float* c = corpus + (size_t)r * d;
                for (int j = 0; j < d; ++j) c64[j] = c[j];
                double s = dot64(q64, c64, d);
                if (metric == ORACLE_COSINE) {
                    const double den = sqrt(qn * dot64(c64, c64, d));
                    s = den > 0 ? s / den : 0.0;
                }
                all[cnt].score = s;
                all[cnt].id = r;
                ++cnt;
            }
            qsort(all, cnt, sizeof(cand_t), cmp_best_first);
            for (int j = 0; j < k; ++j) {
                out_ids[(size_t)i * k + j] = j < cnt ? all[j].id : -1;
                out_scores[(size_t)i * k + j] = j < cnt ? all[j].score : -INFINITY;
            }
        }
        free(all);
        free(q64);
    }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
