/* TEST INFRASTRUCTURE. C restatement of the reference's correlation generators for the closed-form Matern
 * branches, in the reference's own operation order so that results are bit-identical to its compiled Cython on
 * x86-64 (build with -O2 -ffp-contract=off; glibc exp/sqrt/pow as the reference links):
 *   matern_kernel / euclidean_distance   gaussian_proc/generate_correlation/_kernels.pyx:17-100, :107-136
 *   dense fill (i, j>=i, mirror)         _generate_dense_correlation.pyx:76-91
 *   sparse keep-rule  K_ij > tau         _generate_sparse_correlation.pyx:140-197
 * The general-nu (Bessel) branch lives in oracle/matern.py (scipy.special.kv/gamma, as the reference calls them).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

double oracle_matern_kernel(double x, double nu) {
    if (x == 0) return 1.0;
    if (nu == 0.5) return exp(-x);
    if (nu == 1.5) return (1.0 + sqrt(3.0) * x) * exp(-sqrt(3.0) * x);
    if (nu == 2.5) return (1.0 + sqrt(5.0) * x + (5.0 / 3.0) * pow(x, 2.0)) * exp(-sqrt(5.0) * x);
    if (nu < 100) return NAN; /* Bessel branch: not here */
    return exp(-0.5 * pow(x, 2.0));
}

double oracle_distance(const double* p1, const double* p2, const double* scale, int d) {
    double s = 0;
    for (int k = 0; k < d; ++k) s += pow((p1[k] - p2[k]) / scale[k], 2.0);
    return sqrt(s);
}

void oracle_dense(const double* pts, int64_t n, int d, const double* scale, double nu, double* K) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = i; j < n; ++j) {
            double v = oracle_matern_kernel(oracle_distance(pts + i * d, pts + j * d, scale, d), nu);
            K[i * n + j] = v;
            if (i != j) K[j * n + i] = v;
        }
}

/* Row-wise count / fill of {j : K_ij > tau} in increasing j == the canonical CSR that coo_matrix(...).tocsr()
 * produces from the reference's (i,j),(j,i) insertions (_generate_sparse_correlation.pyx:580-584).
 * Brute force over all pairs, like the reference. pass 0: counts[i]; pass 1: fill using indptr. */
void oracle_sparse_rows(const double* pts, int64_t n, int d, const double* scale, double nu, double tau, int pass,
                        int64_t row_begin, int64_t row_end, int64_t* counts, const int64_t* indptr, int32_t* indices,
                        double* data) {
    for (int64_t i = row_begin; i < row_end; ++i) {
        int64_t c = 0;
        for (int64_t j = 0; j < n; ++j) {
            /* the reference evaluates the pair once with (min, max) ordering: points[i] - points[j] for i <= j */
            const double* a = (i <= j) ? pts + i * d : pts + j * d;
            const double* b = (i <= j) ? pts + j * d : pts + i * d;
            double v = oracle_matern_kernel(oracle_distance(a, b, scale, d), nu);
            if (v > tau) {
                if (pass) { indices[indptr[i] + c] = (int32_t)j; data[indptr[i] + c] = v; }
                ++c;
            }
        }
        if (!pass) counts[i] = c;
    }
}
