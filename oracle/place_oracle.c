/* place_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, plain C).
 *
 * Restates map_mutations_gt + _normalize_log_probs of the reference
 * (cbg-ethz/scSomMerClock, run_PT_test.py:386-448 and :368-383) as dense loops:
 * for every SNV the log-likelihood of sitting on each row of the clade matrix S
 * is the sum over observed cells of log(1-FN) / log(FP) / log(1-FP) / log(FN)
 * (run_PT_test.py:408-427), normalised with the max-shifted log-sum-exp of
 * :368-383 (including its clip at -100) and summed over SNVs (:441).
 * Used by tests (cross-check of the NumPy oracle) and by bench.py's CPU baseline
 * legs; never by the product.  Pinned through tests/test_oracle_golden.py.
 *
 * codes: [n_snv][n_cells] uint8 in {0 wild type, 1 het, 2 hom, 3 missing}
 * S    : [n_rows][n_cells] uint8 0/1
 * soft : [n_rows] doubles (output, overwritten)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int place_oracle(const uint8_t* codes, int64_t n_snv, int32_t n_cells, const uint8_t* S, int32_t n_rows, double FP,
                 double FN, double* soft, int32_t n_threads) {
    const double lTP = log(1.0 - FN), lFP = log(FP), lTN = log(1.0 - FP), lFN = log(FN);
    const double log_min = -100.0, under = -708.3964185322641; /* exp() underflow trap of np.seterr('raise') */
    memset(soft, 0, sizeof(double) * (size_t)n_rows);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    int failed = 0;
#pragma omp parallel
    {
        double* probs = (double*)malloc(sizeof(double) * (size_t)n_rows);
        double* acc = (double*)calloc((size_t)n_rows, sizeof(double));
        if (!probs || !acc) {
            failed = 1;
        } else {
#pragma omp for schedule(dynamic, 16)
            for (int64_t i = 0; i < n_snv; ++i) {
                const uint8_t* g = codes + (size_t)i * n_cells;
                for (int b = 0; b < n_rows; ++b) {
                    const uint8_t* s = S + (size_t)b * n_cells;
                    double p = 0.0;
                    for (int c = 0; c < n_cells; ++c) {
                        if (g[c] == 3) continue;               /* nansum */
                        if (g[c] != 0) p += s[c] ? lTP : lFP;   /* .astype(bool): het and hom alike */
                        else p += s[c] ? lFN : lTN;
                    }
                    probs[b] = p;
                }
                int mx = 0;
                for (int b = 1; b < n_rows; ++b)
                    if (probs[b] > probs[mx]) mx = b;           /* nanargmax: first maximum */
                int clip = 0;
                for (int b = 0; b < n_rows; ++b)
                    if (b != mx && probs[b] - probs[mx] < under) clip = 1;
                double se = 0.0;
                for (int b = 0; b < n_rows; ++b) {
                    if (b == mx) continue;
                    double d = probs[b] - probs[mx];
                    if (clip && d < log_min) d = log_min;
                    se += exp(d);
                }
                const double l1p = log1p(se);
                for (int b = 0; b < n_rows; ++b) {
                    double pn = probs[b] - probs[mx] - l1p;
                    if (pn < log_min) pn = log_min;
                    if (pn > 0.0) pn = 0.0;
                    acc[b] += exp(pn);
                }
            }
#pragma omp critical
            for (int b = 0; b < n_rows; ++b) soft[b] += acc[b];
        }
        free(probs);
        free(acc);
    }
    return failed;
}
