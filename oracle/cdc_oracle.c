/* TEST INFRASTRUCTURE — CPU oracle, never the product and never on the product path.
 *
 * Plain-C restatement of the arithmetic of the reference hot path (jrhaberstroh/cGromCorrFMO):
 *   ComputeCDC_v1            reference src/grom2atomFMO.cpp:625-739
 *   unit conversion          reference src/FMOparse.cpp:151-154
 *   Correlate + finalise     reference src/FMOparse.cpp:168-178, 188-206
 * Compiled with -ffp-contract=off so the float32 variants perform exactly the reference's operations in the reference's
 * order (g++ -O2 on x86-64 does not contract either).  Pinned: oracle_cdc_f32 is checked BIT-EXACT against the
 * unmodified reference (oracle/_ref/ref_harness) in tests/test_oracle.py and against tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* float32, sequential — the reference's own arithmetic.
 * groups[a] < 0: atom of chromophore -groups[a]; > 0: atom of environment group groups[a].
 * q[a] is AtomDataLookup_v2's output for this site: dQ for the site's atoms, table charge for everyone else.
 * Column = g-1 for chromophore g, nChromo+g-1 for environment group g (reference :688-693).  The reference's outer loop
 * runs i = 1..nGroups (:669), so any chromophore index > nGroups is never visited; kept.
 * A single pass over atoms adds to each column in the same order as the reference's (group, atom, site-atom) nest. */
void oracle_cdc_f32(int site, int natoms, const float *Xi, const float *q, const int *groups,
                    int nGroups, int nChromo, float *cdc_kcal)
{
    int numTot = nGroups + nChromo;
    int *cidx = (int *)malloc(sizeof(int) * (size_t)natoms);
    int nc = 0;
    for (int a = 0; a < natoms; a++) if (groups[a] == -site) cidx[nc++] = a;
    for (int i = 0; i < numTot; i++) cdc_kcal[i] = 0.0f;
    float esConst = 33.2f;
    for (int a = 0; a < natoms; a++) {
        int g = groups[a];
        int ag = abs(g);
        if (ag < 1 || ag > nGroups || -g == site) continue;
        int col = g < 0 ? ag - 1 : nChromo + ag - 1;
        for (int c = 0; c < nc; c++) {
            int ci = cidx[c];
            float d2 = 0.f;
            for (int dir = 0; dir < 3; dir++) {
                float dx = Xi[3 * a + dir] - Xi[3 * ci + dir];
                d2 += dx * dx;
            }
            float d = sqrtf(d2);
            float coupling = esConst / d * (q[a] * q[ci]);
            cdc_kcal[col] += coupling;
        }
    }
    free(cidx);
}

/* float64 restatement on the SAME float32 inputs: every operation in double, double accumulation.
 * scale[col] = sum over pairs of |term| (the cancellation-aware tolerance scale). */
void oracle_cdc_f64(int site, int natoms, const float *Xi, const float *q, const int *groups,
                    int nGroups, int nChromo, double *cdc_kcal, double *scale)
{
    int numTot = nGroups + nChromo;
    int *cidx = (int *)malloc(sizeof(int) * (size_t)natoms);
    int nc = 0;
    for (int a = 0; a < natoms; a++) if (groups[a] == -site) cidx[nc++] = a;
    for (int i = 0; i < numTot; i++) { cdc_kcal[i] = 0.0; scale[i] = 0.0; }
    double esConst = (double)33.2f;
    for (int a = 0; a < natoms; a++) {
        int g = groups[a];
        int ag = abs(g);
        if (ag < 1 || ag > nGroups || -g == site) continue;
        int col = g < 0 ? ag - 1 : nChromo + ag - 1;
        double acc = 0.0, sc = 0.0;
        for (int c = 0; c < nc; c++) {
            int ci = cidx[c];
            double d2 = 0.0;
            for (int dir = 0; dir < 3; dir++) {
                double dx = (double)Xi[3 * a + dir] - (double)Xi[3 * ci + dir];
                d2 += dx * dx;
            }
            double t = esConst / sqrt(d2) * ((double)q[a] * (double)q[ci]);
            acc += t;
            sc += fabs(t);
        }
        cdc_kcal[col] += acc;
        scale[col] += sc;
    }
    free(cidx);
}

/* reference src/FMOparse.cpp:151-154: cdc_cm = float( double(cdc_kcal) * 349.7 ) */
void oracle_kcal_to_cm(int n, const float *kcal, float *cm)
{
    for (int i = 0; i < n; i++) cm[i] = (float)((double)kcal[i] * 349.7);
}

/* reference src/FMOparse.cpp:171-177: one rank-1 float32 update of (sum, sumsq) for one (frame, site) row */
void oracle_correlate_accum_f32(int numTot, const float *row_kcal, float *sum, float *sumsq)
{
    for (int i = 0; i < numTot; i++) {
        sum[i] += row_kcal[i];
        for (int j = 0; j < numTot; j++) sumsq[(size_t)numTot * i + j] += row_kcal[i] * row_kcal[j];
    }
}

/* reference src/FMOparse.cpp:189-206, LITERALLY, including the stride variable dEsize that stays 0 (:42):
 * row 0 of sumsq is divided by T numTot times, rows >= 1 are neither normalised nor read, and
 * mtx[i][j] = sumsq[j] - mean[i]*mean[j].  sum/sumsq are modified in place like the reference does. */
void oracle_finalise_mtx_compat_f32(int numTot, int T, float *sum, float *sumsq, float *mtx)
{
    const int dEsize = 0;
    for (int i = 0; i < numTot; i++) {
        sum[i] /= T;
        for (int j = 0; j < numTot; j++) sumsq[(size_t)(dEsize * i) + j] /= T;
    }
    for (int i = 0; i < numTot; i++)
        for (int j = 0; j < numTot; j++)
            mtx[(size_t)numTot * i + j] = sumsq[(size_t)(dEsize * i) + j] - sum[i] * sum[j];
}

/* The formula the reference's structure intends: population covariance in double,
 * E[x_i x_j] - E[x_i]E[x_j] computed as a centred two-pass sum from the float32 dE matrix X[T][numTot]. */
void oracle_cov_f64(int T, int numTot, const float *X, int ddof, double *mean, double *cov)
{
    for (int i = 0; i < numTot; i++) mean[i] = 0.0;
    for (int t = 0; t < T; t++)
        for (int i = 0; i < numTot; i++) mean[i] += (double)X[(size_t)t * numTot + i];
    for (int i = 0; i < numTot; i++) mean[i] /= T;
    memset(cov, 0, sizeof(double) * (size_t)numTot * numTot);
    double *d = (double *)malloc(sizeof(double) * (size_t)numTot);
    for (int t = 0; t < T; t++) {
        for (int i = 0; i < numTot; i++) d[i] = (double)X[(size_t)t * numTot + i] - mean[i];
        for (int i = 0; i < numTot; i++) {
            double di = d[i];
            if (di == 0.0) continue;
            double *row = cov + (size_t)numTot * i;
            for (int j = 0; j < numTot; j++) row[j] += di * d[j];
        }
    }
    double div = (double)(T - ddof);
    for (size_t k = 0; k < (size_t)numTot * numTot; k++) cov[k] /= div;
    free(d);
}

/* Checker for K1's division-free "%8.3f" conversion: for every integer n in [n_lo, n_hi) (n < 2^24, so (float)n is
 * exact), compare q + (n - q*1000)*r, q = n*r, r = RN(1/1000) — two fused multiply-adds — with the IEEE quotient
 * (float)n / 1000.0f, which equals the reference's (float)atof("%d.%03d") (src/grom2atomFMO.cpp:448-450).
 * Returns the number of mismatching bit patterns. */
long cgo_div1000_mismatches(unsigned n_lo, unsigned n_hi) {
  const float r = 1.0f / 1000.0f;
  long bad = 0;
  for (unsigned n = n_lo; n < n_hi; n++) {
    const float nf = (float)n;
    const float q = nf * r;
    const float q2 = fmaf(fmaf(-q, 1000.0f, nf), r, q);
    const float ref = nf / 1000.0f;
    if (memcmp(&q2, &ref, sizeof(float))) bad++;
  }
  return bad;
}

/* The same quotient against the reference's literal text path: snprintf("%u.%03u") -> atof -> (float). */
long cgo_div1000_vs_atof_mismatches(unsigned n_lo, unsigned n_hi, unsigned step) {
  const float r = 1.0f / 1000.0f;
  long bad = 0;
  char buf[32];
  for (unsigned n = n_lo; n < n_hi; n += step) {
    const float nf = (float)n;
    const float q = nf * r;
    const float q2 = fmaf(fmaf(-q, 1000.0f, nf), r, q);
    snprintf(buf, sizeof buf, "%u.%03u", n / 1000u, n % 1000u);
    const float ref = (float)atof(buf);
    if (memcmp(&q2, &ref, sizeof(float))) bad++;
  }
  return bad;
}

/* Checker for K5 (the `.time` text formatted on the device): rows x ncol float32 printed the way the reference prints them,
 * `stream << float` == printf("%g") (src/FMOparse.cpp:158-165), ',' between values and '\n' after each row.
 * Returns the number of bytes written to out (the caller provides rows * ncol * 16 bytes). */
long cgo_format_g_rows(const float *v, long rows, int ncol, char *out) {
  char *p = out;
  for (long r = 0; r < rows; r++) {
    for (int j = 0; j < ncol; j++) {
      p += sprintf(p, "%g", (double)v[r * ncol + j]);
      *p++ = (j == ncol - 1) ? '\n' : ',';
    }
  }
  return (long)(p - out);
}
