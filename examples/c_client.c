/* c_client.c -- a pure-C client of libcmwdense.so: no CUDA headers, no Python, no torch.
 *
 * Shows the drop-in boundary of include/cmw_dense.h the way a non-Python host (or a ctypes / cgo /
 * JNI stub) would use it: create a store, append rows from host memory, search from host memory.
 * It does what ChromaStore.add_async + similarity_search_async do in the reference
 * (rag_engine/storage/vector_store.py:54-82), and checks the answer against a brute-force loop.
 *
 *   gcc -O2 -Iinclude examples/c_client.c -o /tmp/c_client -Lcmw_rag_b200/csrc -lcmwdense \
 *       -Wl,-rpath,$PWD/cmw_rag_b200/csrc -lm && /tmp/c_client
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cmw_dense.h"

static unsigned long long rng_state = 88172645463325252ull;
static double rnd(void) { /* xorshift64 -> uniform (0,1) */
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (double)(rng_state >> 11) / 9007199254740992.0 + 1e-18;
}
static float gauss(void) { return (float)(sqrt(-2.0 * log(rnd())) * cos(6.283185307179586 * rnd())); }

int main(void) {
    const int n = 20000, d = 384, b = 5, k = 10;
    float* rows = malloc(sizeof(float) * (size_t)n * d);
    float* q = malloc(sizeof(float) * (size_t)b * d);
    for (size_t i = 0; i < (size_t)n * d; ++i) rows[i] = gauss();
    for (int i = 0; i < b; ++i) /* queries near rows 100*i+3 */
        for (int j = 0; j < d; ++j) q[(size_t)i * d + j] = rows[(size_t)(100 * i + 3) * d + j] + 0.3f * gauss();

    cmw_store* store = NULL;
    if (cmw_store_create(0, d, n, CMW_STORE_F32 | CMW_STORE_BF16, 0, &store) != 0) {
        fprintf(stderr, "create failed: %s\n", cmw_last_error());
        return 2;
    }
    if (cmw_store_append_host_f32(store, rows, NULL, n) != 0) {
        fprintf(stderr, "append failed: %s\n", cmw_last_error());
        return 2;
    }
    float* scores = malloc(sizeof(float) * b * k);
    int64_t* ids = malloc(sizeof(int64_t) * b * k);
    int32_t flags[5];
    if (cmw_search_host(store, q, b, k, CMW_METRIC_COSINE, CMW_MODE_F32_EXACT, scores, ids, flags) != 0) {
        fprintf(stderr, "search failed: %s\n", cmw_last_error());
        return 2;
    }
    /* brute-force check of the best hit of every query (fp64 cosine) */
    int bad = 0;
    for (int i = 0; i < b; ++i) {
        double best = -2.0, qn = 0.0;
        long best_id = -1;
        for (int j = 0; j < d; ++j) qn += (double)q[(size_t)i * d + j] * q[(size_t)i * d + j];
        for (int r = 0; r < n; ++r) {
            double dot = 0.0, cn = 0.0;
            for (int j = 0; j < d; ++j) {
                dot += (double)q[(size_t)i * d + j] * rows[(size_t)r * d + j];
                cn += (double)rows[(size_t)r * d + j] * rows[(size_t)r * d + j];
            }
            const double c = dot / sqrt(qn * cn);
            if (c > best) { best = c; best_id = r; }
        }
        printf("query %d: top-1 id %lld score %.6f (brute force: id %ld score %.6f) flags %d\n", i,
               (long long)ids[i * k], scores[i * k], best_id, best, flags[i]);
        if (ids[i * k] != best_id || fabs(scores[i * k] - best) > 1e-5 || best_id != 100 * i + 3 || flags[i] != 0) bad++;
        for (int j = 1; j < k; ++j)
            if (scores[i * k + j] > scores[i * k + j - 1]) bad++;
    }
    cmw_store_info info;
    cmw_store_get_info(store, &info);
    printf("rows %lld, hbm %.1f MB, kernels launched %lld, abi %d\n", (long long)info.rows, info.hbm_bytes / 1e6,
           (long long)cmw_kernel_launches(), cmw_abi_version());
    cmw_store_destroy(store);
    free(rows); free(q); free(scores); free(ids);
    if (bad) { fprintf(stderr, "MISMATCH (%d)\n", bad); return 1; }
    printf("c_client ok\n");
    return 0;
}
