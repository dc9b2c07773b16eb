/* Plain-C restatement of the exact kNN the reference approximates with FLANN
 * (TEST INFRASTRUCTURE ONLY - see oracle/rf_oracle.py header).
 *
 * Reference call site: util/retrieval.py:92
 *     results, dists = flann_obj.nn_index(feature_subset, 2 * K, checks=...)
 * pyflann is un-vendored, unpinned and approximate, so parity for this piece is
 * UNPINNED; this file defines the canonical exact rule instead:
 *
 *   d(q,x) = sum over i = 0..D-1 (in that order) of (double(q_i)-double(x_i))^2,
 *            every operation correctly rounded in binary64, NO fma contraction
 *            (compile with -ffp-contract=off);
 *   result = the k rows smallest under ascending (d, row index);
 *   dist   = (float)d.
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC (oracle/Makefile).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double d; int i; } cand_t;

static inline int cand_less(double d, int i, const cand_t *c) {
    return d < c->d || (d == c->d && i < c->i);
}

typedef struct {
    const float *db; long N; const float *q; long Q; int D; int k;
    int *out_idx; float *out_dist; long q_begin, q_end; int fail;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    const int D = j->D, k = j->k;
    cand_t *best = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
    double *qd = (double *)malloc(sizeof(double) * (size_t)D);
    if (!best || !qd) { j->fail = 1; free(best); free(qd); return 0; }
    for (long qi = j->q_begin; qi < j->q_end; ++qi) {
        for (int i = 0; i < D; ++i) qd[i] = (double)j->q[qi * D + i];
        int n = 0;
        for (long r = 0; r < j->N; ++r) {
            const float *x = j->db + r * D;
            double acc = 0.0;
            for (int i = 0; i < D; ++i) {
                double diff = qd[i] - (double)x[i];
                double sq = diff * diff;
                acc = acc + sq;
            }
            if (n < k) {
                int p = n++;
                while (p > 0 && cand_less(acc, (int)r, &best[p - 1])) { best[p] = best[p - 1]; --p; }
                best[p].d = acc; best[p].i = (int)r;
            } else if (cand_less(acc, (int)r, &best[k - 1])) {
                int p = k - 1;
                while (p > 0 && cand_less(acc, (int)r, &best[p - 1])) { best[p] = best[p - 1]; --p; }
                best[p].d = acc; best[p].i = (int)r;
            }
        }
        for (int t = 0; t < k; ++t) {
            j->out_idx[qi * k + t] = best[t].i;
            j->out_dist[qi * k + t] = (float)best[t].d;
        }
    }
    free(best); free(qd);
    return 0;
}

int rf_oracle_knn(const float *db, long N, const float *q, long Q, int D, int k,
                  int *out_idx, float *out_dist, int threads) {
    if (k <= 0 || k > N || D <= 0 || D > 4096) return 1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > Q) threads = Q > 0 ? (int)Q : 1;
    pthread_t tid[256];
    job_t jobs[256];
    long per = (Q + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        long b = t * per, e = b + per; if (e > Q) e = Q; if (b > Q) b = Q;
        job_t jb = { db, N, q, Q, D, k, out_idx, out_dist, b, e, 0 };
        jobs[t] = jb;
        pthread_create(&tid[t], 0, worker, &jobs[t]);
    }
    int fail = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(tid[t], 0); fail |= jobs[t].fail; }
    return fail;
}

/* fp32 brute force "what torch.cdist+topk would do" is NOT provided here on
 * purpose: the CPU baseline for timing uses torch (bench.py), the oracle is
 * only ever the checker. */
