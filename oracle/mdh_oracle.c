/*
 * mdh_oracle.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  Nothing under mdhelper_b200/
 * imports it; the product path has no CPU fallback.
 *
 * What it restates (plain C, fp64, no FMA contraction: build with
 * -O2 -ffp-contract=off, never -ffast-math):
 *
 *  (1) The distance arithmetic behind the reference's radial_histogram
 *      (/root/reference/src/mdhelper/analysis/structure.py:93-96), which is the
 *      THIRD-PARTY call MDAnalysis.lib.distances.capped_distance.  MDAnalysis
 *      (requirements.txt:3 "mdanalysis>=2.2.0", no exact pin, not vendored) is
 *      absent from /root/reference and from this image, so this is a
 *      restatement of its published algorithm:
 *        dx_k  = (double)(float)(pos2[j][k] - pos1[i][k])
 *        inv_k = (float)(1.0 / box_k)
 *        s     = (double)inv_k * dx_k
 *        dx_k  = (double)box_k * (s - round(s))          (C round())
 *        d     = sqrt((dx0*dx0 + dx1*dx1) + dx2*dx2)      (products rounded separately)
 *        keep iff  min_cutoff < d <= max_cutoff
 *      PARITY STATUS: "parity unpinned" for this piece against real MDAnalysis
 *      (it cannot be run here); it is pinned against the reference's own
 *      known-answer construction tests/test_analysis_structure.py:21-40 and the
 *      binning is the real numpy.histogram (see oracle/reference_port.py).
 *
 *  (2) The direct-sum Fourier kernel of the reference,
 *      /root/reference/src/mdhelper/algorithm/accelerated.py:81-122 (serial) and
 *      :124-165 (prange over wavevectors):  F[i] = sum_j exp(i q_i . r_j).
 *      PARITY STATUS: pinned -- checked in this container against the real numba
 *      kernels (tests/golden/make_golden.py; fixtures under tests/golden/).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- (1) minimum-image distances ------------------------------------------------ */

/* One pair.  Follows SURVEY.md Appendix A item 3 to the letter. */
static inline double min_image_dist2(const float *a /*pos1[i]*/, const float *b /*pos2[j]*/,
                                     const float box[3], const float inv[3])
{
    double dx[3];
    for (int k = 0; k < 3; ++k) {
        float df = b[k] - a[k];              /* float32 subtraction */
        double d = (double)df;
        if (box[k] > FLT_EPSILON) {
            double s = (double)inv[k] * d;
            d = (double)box[k] * (s - round(s));
        }
        dx[k] = d;
    }
    return (dx[0] * dx[0] + dx[1] * dx[1]) + dx[2] * dx[2];
}

static inline void inverse_box(const float box[3], float inv[3])
{
    for (int k = 0; k < 3; ++k) inv[k] = (float)(1.0 / (double)box[k]);
}

/*
 * Restated MDAnalysis `_ortho_pbc` (lib/include/calc_distances.h, reached through
 * apply_PBC from the grid search `FastNS`, lib/nsgrid.pyx) [recall; "parity unpinned"]:
 * moves coordinates into the primary cell IN FLOAT32 STORAGE.  One box shift is tried
 * first (computed in double, stored as float); only if that is not enough the number of
 * shifts is estimated with floor() and applied in float, followed by one corrective
 * single shift.  A coordinate just below 0 can end up exactly on the box edge (the sum
 * rounds to box_k), as in MDAnalysis.
 */
void mdho_ortho_pbc(float *coords, int64_t n, const float *box)
{
    if (!box[0] && !box[1] && !box[2]) return;
    double inverse_box[3];
    for (int j = 0; j < 3; ++j)
        inverse_box[j] = box[j] > FLT_EPSILON ? 1.0 / (double)box[j] : 0.0;
    for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < 3; ++j) {
            float *c = coords + 3 * i + j;
            double crd = (double)*c;
            if (crd < 0.0) {
                crd += box[j];
                if (crd < 0.0) {                   /* more than one box away */
                    int s = (int)floor(*c * inverse_box[j]);
                    *c -= s * box[j];
                    if (*c < 0.0) *c += box[j];
                } else {
                    *c = (float)crd;
                }
            }
            /* no "else": a single shift up may have produced exactly box_k */
            if (crd >= box[j]) {
                crd -= box[j];
                if (crd >= box[j]) {
                    int s = (int)floor(*c * inverse_box[j]);
                    *c -= s * box[j];
                    if (*c >= box[j]) *c -= box[j];
                } else {
                    *c = (float)crd;
                }
            }
        }
}

/*
 * Brute-force capped distances over rows [i0, i1) of pos1 against all of pos2.
 * Writes up to cap results; returns the number of pairs that satisfy the
 * cut-offs (which may exceed cap: call again with a larger buffer).
 * pairs may be NULL (distances only).
 */
int64_t mdho_capped_distance_bruteforce(const float *pos1, int64_t i0, int64_t i1,
                                        const float *pos2, int64_t n2, const float *box,
                                        double max_cutoff, double min_cutoff,
                                        int64_t *pairs, double *dist, int64_t cap)
{
    float inv[3];
    inverse_box(box, inv);
    int64_t m = 0;
    for (int64_t i = i0; i < i1; ++i) {
        const float *a = pos1 + 3 * i;
        for (int64_t j = 0; j < n2; ++j) {
            double d = sqrt(min_image_dist2(a, pos2 + 3 * j, box, inv));
            if (d > min_cutoff && d <= max_cutoff) {
                if (m < cap) {
                    if (pairs) { pairs[2 * m] = i; pairs[2 * m + 1] = j; }
                    dist[m] = d;
                }
                ++m;
            }
        }
    }
    return m;
}

/*
 * Cell-list (grid search) capped distances: pos2 is binned into cells of edge
 * >= max_cutoff, each pos1[i] visits the 27 surrounding cells.  The per-pair
 * arithmetic is the same function as the brute-force path, so for coordinates
 * inside [0, L) the two agree bit for bit.
 * Requires >= 3 cells per axis; returns -1 if the box is too small for that.
 */
int64_t mdho_capped_distance_cells(const float *pos1, int64_t n1, const float *pos2,
                                   int64_t n2, const float *box, double max_cutoff,
                                   double min_cutoff, int64_t *pairs, double *dist,
                                   int64_t cap)
{
    float inv[3];
    inverse_box(box, inv);
    int nc[3];
    double cs[3];
    for (int k = 0; k < 3; ++k) {
        nc[k] = (int)floor((double)box[k] / (max_cutoff * 1.00001));
        if (nc[k] < 3) return -1;
        if (nc[k] > 1024) nc[k] = 1024;
        cs[k] = (double)box[k] / nc[k];
    }
    int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    int64_t *start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
    int64_t *cell_of = (int64_t *)malloc((size_t)(n2 > 0 ? n2 : 1) * sizeof(int64_t));
    int64_t *order = (int64_t *)malloc((size_t)(n2 > 0 ? n2 : 1) * sizeof(int64_t));
    if (!start || !cell_of || !order) { free(start); free(cell_of); free(order); return -2; }

#define CELL_COORD(x, k, out)                                   \
    do {                                                        \
        double w_ = (double)(x) - floor((double)(x) / (double)box[k]) * (double)box[k]; \
        int c_ = (int)floor(w_ / cs[k]);                        \
        if (c_ < 0) c_ = 0;                                     \
        if (c_ >= nc[k]) c_ = nc[k] - 1;                        \
        (out) = c_;                                             \
    } while (0)

    for (int64_t j = 0; j < n2; ++j) {
        int c[3];
        for (int k = 0; k < 3; ++k) CELL_COORD(pos2[3 * j + k], k, c[k]);
        cell_of[j] = ((int64_t)c[2] * nc[1] + c[1]) * nc[0] + c[0];
        start[cell_of[j] + 1]++;
    }
    for (int64_t c = 0; c < ncell; ++c) start[c + 1] += start[c];
    {
        int64_t *fill = (int64_t *)malloc((size_t)ncell * sizeof(int64_t));
        if (!fill) { free(start); free(cell_of); free(order); return -2; }
        memcpy(fill, start, (size_t)ncell * sizeof(int64_t));
        for (int64_t j = 0; j < n2; ++j) order[fill[cell_of[j]]++] = j;
        free(fill);
    }

    int64_t m = 0;
    for (int64_t i = 0; i < n1; ++i) {
        const float *a = pos1 + 3 * i;
        int c[3];
        for (int k = 0; k < 3; ++k) CELL_COORD(a[k], k, c[k]);
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dxc = -1; dxc <= 1; ++dxc) {
                    int cx = (c[0] + dxc + nc[0]) % nc[0];
                    int cy = (c[1] + dy + nc[1]) % nc[1];
                    int cz = (c[2] + dz + nc[2]) % nc[2];
                    int64_t cc = ((int64_t)cz * nc[1] + cy) * nc[0] + cx;
                    for (int64_t p = start[cc]; p < start[cc + 1]; ++p) {
                        int64_t j = order[p];
                        double d = sqrt(min_image_dist2(a, pos2 + 3 * j, box, inv));
                        if (d > min_cutoff && d <= max_cutoff) {
                            if (m < cap) {
                                if (pairs) { pairs[2 * m] = i; pairs[2 * m + 1] = j; }
                                dist[m] = d;
                            }
                            ++m;
                        }
                    }
                }
    }
#undef CELL_COORD
    free(start); free(cell_of); free(order);
    return m;
}

/*
 * Triclinic cells (box9: row-major lower-triangular float32 matrix a_x 0 0 / b_x b_y 0 /
 * c_x c_y c_z, MDAnalysis' triclinic_vectors).  Restates, from the published algorithm
 * of MDAnalysis 2.x lib/include/calc_distances.h ("parity unpinned", as above):
 *   _triclinic_pbc      both coordinate sets are moved into the primary cell first.
 *                       Restated mathematically (c, then b, then a: s = floor(x_k / h_kk),
 *                       r -= s h_k in double, rounded to float32); the identity for
 *                       coordinates that already lie in the cell.
 *   minimum_image_triclinic   the shortest of the 27 images dx + ix a + iy b + iz c,
 *                       loop order ix, iy, iz, strict "<", sums in double.
 */
static void triclinic_wrap(const float *in, float *out, int64_t n, const float *h)
{
    for (int64_t i = 0; i < n; ++i) {
        double x = in[3 * i], y = in[3 * i + 1], z = in[3 * i + 2];
        double s = floor(z / (double)h[8]);
        if (s != 0.0) { x -= s * (double)h[6]; y -= s * (double)h[7]; z -= s * (double)h[8]; }
        s = floor(y / (double)h[4]);
        if (s != 0.0) { x -= s * (double)h[3]; y -= s * (double)h[4]; }
        s = floor(x / (double)h[0]);
        if (s != 0.0) x -= s * (double)h[0];
        out[3 * i] = (float)x; out[3 * i + 1] = (float)y; out[3 * i + 2] = (float)z;
    }
}

static inline double triclinic_dist2(const float *a, const float *b, const float *h)
{
    double dx[3];
    for (int k = 0; k < 3; ++k) dx[k] = (double)(float)(b[k] - a[k]);
    double best = (double)FLT_MAX;
    for (int ix = -1; ix < 2; ++ix) {
        double rx = dx[0] + (double)(h[0] * ix);
        for (int iy = -1; iy < 2; ++iy) {
            double ry0 = rx + (double)(h[3] * iy);
            double ry1 = dx[1] + (double)(h[4] * iy);
            for (int iz = -1; iz < 2; ++iz) {
                double rz0 = ry0 + (double)(h[6] * iz);
                double rz1 = ry1 + (double)(h[7] * iz);
                double rz2 = dx[2] + (double)(h[8] * iz);
                double dsq = (rz0 * rz0 + rz1 * rz1) + rz2 * rz2;
                if (dsq < best) best = dsq;
            }
        }
    }
    return best;
}

int64_t mdho_capped_distance_triclinic(const float *pos1, int64_t n1, const float *pos2,
                                       int64_t n2, const float *box9, double max_cutoff,
                                       double min_cutoff, int64_t *pairs, double *dist,
                                       int64_t cap)
{
    float *w1 = (float *)malloc(sizeof(float) * 3 * (size_t)(n1 > 0 ? n1 : 1));
    float *w2 = (float *)malloc(sizeof(float) * 3 * (size_t)(n2 > 0 ? n2 : 1));
    if (!w1 || !w2) { free(w1); free(w2); return -2; }
    triclinic_wrap(pos1, w1, n1, box9);
    triclinic_wrap(pos2, w2, n2, box9);
    int64_t m = 0;
    for (int64_t i = 0; i < n1; ++i)
        for (int64_t j = 0; j < n2; ++j) {
            double d = sqrt(triclinic_dist2(w1 + 3 * i, w2 + 3 * j, box9));
            if (d > min_cutoff && d <= max_cutoff) {
                if (m < cap) {
                    if (pairs) { pairs[2 * m] = i; pairs[2 * m + 1] = j; }
                    dist[m] = d;
                }
                ++m;
            }
        }
    free(w1); free(w2);
    return m;
}

/* ---- (2) direct-sum Fourier transform of delta functions --------------------- */

/*
 * F[i] = sum_j exp(i * (qs[i] . rs[j])); out is interleaved (re, im).
 * Mirrors accelerated.py:117-122 (serial) / :160-165 (prange over i when
 * n_threads > 1).  The reference is numba fastmath=True, i.e. not itself
 * bit-reproducible; agreement is to ~1e-13 relative, far inside the 1e-6 bar.
 */
void mdho_delta_fourier_transform_sum(const double *qs, int64_t nq, const double *rs,
                                      int64_t n, double *out, int n_threads)
{
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(n_threads > 0 ? n_threads : 1)
#endif
    for (int64_t i = 0; i < nq; ++i) {
        const double q0 = qs[3 * i], q1 = qs[3 * i + 1], q2 = qs[3 * i + 2];
        double re = 0.0, im = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            double ph = q0 * rs[3 * j] + q1 * rs[3 * j + 1] + q2 * rs[3 * j + 2];
            re += cos(ph);
            im += sin(ph);
        }
        out[2 * i] = re;
        out[2 * i + 1] = im;
    }
}

int mdho_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
