/* CPU oracle in C -- TEST INFRASTRUCTURE ONLY (see oracle/bruteforce_oracle.py).
 *
 * Float64 restatement of the reference brute force for row samples of problems
 * whose dense matrix cannot exist (N = M = 10^6): one pass over the sources per
 * target row, nothing materialised.  Follows, in
 * /root/reference/kernel_matrix_benchmarks/algorithms/bruteforce.py:
 *   :53-54   squared distance as a sum of squared differences (float64)
 *   :18-22   gaussian exp(-d2), absolute-exponential exp(-sqrt(max(d2,0))),
 *            inverse-distance 1/sqrt(max(d2,0))
 *   :12-14   inverse-distance zeroes flat indices that are multiples of M+1
 *   :142-145 attention = product against [b, 1], then divide
 *   :150     density = row sums
 * Validated against the NumPy oracle (itself pinned by the reference's golden
 * vectors) in tests/test_oracle_c.py.  Threads: pthreads over interleaved
 * target rows (this image's gcc has no libgomp), one per online core.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

enum { KMB_GAUSSIAN = 0, KMB_ABS_EXP = 1, KMB_INV_DIST = 2 };

static int g_threads = 0; /* 0 = one per online core */

int kmb_oracle_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 256 ? 256 : (int)n);
}
void kmb_oracle_set_threads(int n) { g_threads = n; }

typedef struct {
    int kernel_id, D, E, normalize_rows, tid, nthreads;
    const double *x, *y, *b;
    const int64_t* row_ids;
    double* out;
    int64_t n_rows, M;
} job_t;

static void* worker(void* arg) {
    const job_t* J = (const job_t*)arg;
    const int D = J->D, E = J->E;
    const int64_t M = J->M;
    for (int64_t r = J->tid; r < J->n_rows; r += J->nthreads) {
        const double* xr = J->x + r * D;
        const int64_t gi = J->row_ids ? J->row_ids[r] : r;
        const int64_t jz = gi % (M + 1); /* zeroed column of this row, if < M */
        double acc[256];
        double ksum = 0.0;
        for (int e = 0; e < E; ++e) acc[e] = 0.0;
        for (int64_t j = 0; j < M; ++j) {
            const double* yj = J->y + j * D;
            double d2 = 0.0;
            for (int d = 0; d < D; ++d) {
                double t = xr[d] - yj[d];
                d2 += t * t;
            }
            double k;
            if (J->kernel_id == KMB_GAUSSIAN) k = exp(-d2);
            else if (J->kernel_id == KMB_ABS_EXP) k = exp(-sqrt(d2 > 0 ? d2 : 0));
            else k = (j == jz) ? 0.0 : 1.0 / sqrt(d2 > 0 ? d2 : 0);
            ksum += k;
            if (J->b) for (int e = 0; e < E; ++e) acc[e] += k * J->b[j * E + e];
            else acc[0] += k;
        }
        for (int e = 0; e < E; ++e) J->out[r * E + e] = J->normalize_rows ? acc[e] / ksum : acc[e];
    }
    return NULL;
}

/* x: (n_rows, D) rows to evaluate, row_ids: their global indices (NULL = 0..n_rows-1),
 * y: (M, D), b: (M, E) or NULL for density, out: (n_rows, E). Returns 0, or 1 on bad args. */
int kmb_oracle_product_f64(int kernel_id, const double* x, const int64_t* row_ids, const double* y,
                           const double* b, double* out, int64_t n_rows, int64_t M, int D, int E,
                           int normalize_rows) {
    if (kernel_id < 0 || kernel_id > 2 || D < 1 || E < 1 || E > 256) return 1;
    int T = kmb_oracle_threads();
    if (T > n_rows) T = n_rows > 0 ? (int)n_rows : 1;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < T; ++t) {
        job_t j = {kernel_id, D, E, normalize_rows, t, T, x, y, b, row_ids, out, n_rows, M};
        jobs[t] = j;
        if (t > 0 && pthread_create(&th[t], NULL, worker, &jobs[t]) != 0) return 2;
    }
    worker(&jobs[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
    return 0;
}
