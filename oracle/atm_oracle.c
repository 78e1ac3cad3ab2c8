/*
 * atm_oracle.c -- CPU restatement of the ATM Meta-Force hot path (see atm_oracle.h for the status header).
 * TEST INFRASTRUCTURE ONLY.  Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC).
 */
#include "atm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int atm_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py --impl reference: torchrun exports OMP_NUM_THREADS=1, the CPU arm wants every core it may run on */
void atm_oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Scalar stage
 * ---------------------------------------------------------------------------------------------- */

/* Soft-core map u -> u_sc and its derivative fp = du_sc/du.
 * Follows ReferenceATMMetaForceKernels.cpp:26-37 (same function at CommonATMMetaForceKernels.cpp:19-30). */
double atm_oracle_softcore(double u, double umax, double a, double ub, double *fp) {
    if (u <= ub) {
        *fp = 1.0;
        return u;
    }
    double g = (u - ub) / (a * (umax - ub));
    double zeta = 1.0 + 2.0 * g * (g + 1.0);
    double z = pow(zeta, a);
    double s = 4.0 * (2.0 * g + 1.0) / zeta;
    *fp = s * z / pow(1.0 + z, 2);
    return (umax - ub) * (z - 1.0) / (z + 1.0) + ub;
}

/* ReferenceATMMetaForceKernels.cpp:82-98 / CommonATMMetaForceKernels.cpp:182-199. */
void atm_oracle_scalars(const double p[9], double U1, double U2, double out[7]) {
    const double lambda1 = p[0], lambda2 = p[1], alpha = p[2], u0 = p[3], w0 = p[4];
    const double umax = p[5], ubcore = p[6], acore = p[7], direction = p[8];
    double fp;
    double u = direction > 0 ? U2 - U1 : U1 - U2;
    double e0 = direction > 0 ? U1 : U2;
    double usc = atm_oracle_softcore(u, umax, acore, ubcore, &fp);
    double ebias = 0.0;
    double ee = 1.0 + exp(-alpha * (usc - u0));
    if (alpha > 0) ebias = ((lambda2 - lambda1) / alpha) * log(ee);
    ebias += lambda2 * usc + w0;
    double bfp = (lambda2 - lambda1) / ee + lambda1;
    out[0] = usc;
    out[1] = fp;
    out[2] = ebias;
    out[3] = bfp;
    out[4] = e0 + ebias;
    out[5] = direction > 0 ? bfp * fp : 1.0 - bfp * fp;
    out[6] = bfp * fp;
}

/* ------------------------------------------------------------------------------------------------
 * Displacement table and copy-state
 * ---------------------------------------------------------------------------------------------- */

/* CommonATMMetaForceKernels.cpp:83-103: table is float4 of length padded, zero filled, entry s = d[atom_index[s]]
 * rounded to float (the 'particle' field of the force entry is ignored: entry i of the force is atom i). */
void atm_oracle_displ_table(int n, int padded, const int32_t *atom_index, const double *dxyz, float *displ4) {
    memset(displ4, 0, sizeof(float) * 4 * (size_t)padded);
    for (int s = 0; s < n; s++) {
        int a = atom_index ? atom_index[s] : s;
        displ4[4 * s + 0] = (float)dxyz[3 * a + 0];
        displ4[4 * s + 1] = (float)dxyz[3 * a + 1];
        displ4[4 * s + 2] = (float)dxyz[3 * a + 2];
        displ4[4 * s + 3] = 0.0f;
    }
}

/* kernels/atmmetaforce.cc:33-51 with real = float.  posq.w (charge) gets "+ 0".
 * volatile stores keep the compiler from contracting or reassociating the single add. */
void atm_oracle_copy_state_f32(int n, const float *posq, const float *corr, const float *displ4,
                               float *posq1, float *corr1, float *posq2, float *corr2) {
    for (int i = 0; i < n; i++) {
        for (int c = 0; c < 4; c++) {
            posq1[4 * i + c] = posq[4 * i + c];
            float d = (c < 3) ? displ4[4 * i + c] : 0.0f;
            volatile float s = posq[4 * i + c] + d;
            posq2[4 * i + c] = s;
            if (corr) {
                corr1[4 * i + c] = corr[4 * i + c];
                corr2[4 * i + c] = corr[4 * i + c];
            }
        }
    }
}

/* Same kernel with real = double (double-precision mode): the displacement is still the float table. */
void atm_oracle_copy_state_f64(int n, const double *posq, const float *displ4, double *posq1, double *posq2) {
    for (int i = 0; i < n; i++) {
        for (int c = 0; c < 4; c++) {
            posq1[4 * i + c] = posq[4 * i + c];
            double d = (c < 3) ? (double)displ4[4 * i + c] : 0.0;
            posq2[4 * i + c] = posq[4 * i + c] + d;
        }
    }
}

/* ReferenceATMMetaForceKernels.cpp:116-124: pos1 = pos; pos2 = pos + displ, all double. */
void atm_oracle_copy_state_ref(int n, const double *pos, const double *displ, double *pos1, double *pos2) {
    for (int i = 0; i < 3 * n; i++) {
        pos1[i] = pos[i];
        pos2[i] = pos[i] + displ[i];
    }
}

/* ------------------------------------------------------------------------------------------------
 * Merge
 * ---------------------------------------------------------------------------------------------- */

/* ReferenceATMMetaForceKernels.cpp:99-110. */
void atm_oracle_merge_ref(int n, double *force, const double *f1, const double *f2, double sp, double direction) {
    if (direction > 0) {
        for (int i = 0; i < 3 * n; i++) force[i] += sp * f2[i] + (1.0 - sp) * f1[i];
    } else {
        for (int i = 0; i < 3 * n; i++) force[i] += sp * f1[i] + (1.0 - sp) * f2[i];
    }
}

/* kernels/atmmetaforce.cc:8-16 on the SoA int64 buffers (x block, y block, z block, stride = padded); the blend
 * is done in double (the Reference platform's arithmetic) and rounded to nearest instead of going through float. */
void atm_oracle_hybrid_force_i64(int n, int padded, int64_t *force, const int64_t *f1, const int64_t *f2, double sp) {
    for (int c = 0; c < 3; c++)
        for (int i = 0; i < n; i++) {
            size_t k = (size_t)c * padded + i;
            double v = sp * (double)f2[k] + (1.0 - sp) * (double)f1[k];
            force[k] += (int64_t)llrint(v);
        }
}

/* ------------------------------------------------------------------------------------------------
 * Direct-space NonbondedForce
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    int *start; /* n+1 */
    int *list;  /* partners, both directions */
} excl_csr;

static excl_csr build_excl(const atm_oracle_system *s) {
    excl_csr e;
    e.start = (int *)calloc((size_t)s->n + 2, sizeof(int));
    for (int k = 0; k < s->n_excl; k++) {
        e.start[s->excl[2 * k] + 1]++;
        e.start[s->excl[2 * k + 1] + 1]++;
    }
    for (int i = 0; i < s->n; i++) e.start[i + 1] += e.start[i];
    e.list = (int *)malloc(sizeof(int) * (size_t)(e.start[s->n] + 1));
    int *fill = (int *)calloc((size_t)s->n + 1, sizeof(int));
    for (int k = 0; k < s->n_excl; k++) {
        int a = s->excl[2 * k], b = s->excl[2 * k + 1];
        e.list[e.start[a] + fill[a]++] = b;
        e.list[e.start[b] + fill[b]++] = a;
    }
    free(fill);
    return e;
}

static inline int is_excluded(const excl_csr *e, int i, int j) {
    for (int k = e->start[i]; k < e->start[i + 1]; k++)
        if (e->list[k] == j) return 1;
    return 0;
}

static inline double min_image(double d, double L) { return d - L * floor(d / L + 0.5); }

/* One non-excluded pair inside the cutoff: LJ 12-6 with Lorentz-Berthelot combination and Ewald real space.
 * dEdR over r is returned through *fr so that F_i += fr * (r_i - r_j). */
static inline void pair_terms(double r2, double qq, double sig, double eps, double alpha, double *elj, double *ecoul,
                              double *fr) {
    double r = sqrt(r2), inv_r = 1.0 / r;
    double s2 = sig * sig / r2, s6 = s2 * s2 * s2;
    *elj = 4.0 * eps * s6 * (s6 - 1.0);
    double flj = 24.0 * eps * s6 * (2.0 * s6 - 1.0);  /* -r dE/dr */
    double ar = alpha * r;
    double erfc_ar = erfc(ar);
    double pref = ATM_ORACLE_ONE_4PI_EPS0 * qq * inv_r;
    *ecoul = pref * erfc_ar;
    double fcoul = pref * (erfc_ar + ar * exp(-ar * ar) * M_2_SQRTPI); /* -r dE/dr */
    *fr = (flj + fcoul) / r2;
}

double atm_oracle_nb_direct(const atm_oracle_system *s, const double *pos, double *force, double out[4]) {
    const int n = s->n;
    const double rc = s->cutoff, rc2 = rc * rc;
    const double *L = s->box;
    excl_csr ex = build_excl(s);

    /* cell grid (cells >= cutoff); fall back to all-pairs when any dimension has < 3 cells */
    int nc[3];
    int use_cells = 1;
    for (int d = 0; d < 3; d++) {
        nc[d] = (int)floor(L[d] / rc);
        if (nc[d] < 3) use_cells = 0;
    }
    int ncell = use_cells ? nc[0] * nc[1] * nc[2] : 1;
    int *cell_of = (int *)malloc(sizeof(int) * (size_t)n);
    int *cstart = (int *)calloc((size_t)ncell + 1, sizeof(int));
    int *corder = (int *)malloc(sizeof(int) * (size_t)n);
    double *wp = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        int ci[3];
        for (int d = 0; d < 3; d++) {
            double x = pos[3 * i + d] - L[d] * floor(pos[3 * i + d] / L[d]);
            if (x >= L[d]) x -= L[d];
            wp[3 * i + d] = x;
            ci[d] = use_cells ? (int)(x / L[d] * nc[d]) : 0;
            if (use_cells && ci[d] >= nc[d]) ci[d] = nc[d] - 1;
        }
        cell_of[i] = use_cells ? (ci[0] * nc[1] + ci[1]) * nc[2] + ci[2] : 0;
        cstart[cell_of[i] + 1]++;
    }
    for (int c = 0; c < ncell; c++) cstart[c + 1] += cstart[c];
    {
        int *fill = (int *)calloc((size_t)ncell, sizeof(int));
        for (int i = 0; i < n; i++) corder[cstart[cell_of[i]] + fill[cell_of[i]]++] = i;
        free(fill);
    }

    double e_lj = 0.0, e_c = 0.0;
    int nthreads = atm_oracle_num_threads();
    double *fpriv = NULL;
    if (force) fpriv = (double *)calloc((size_t)nthreads * 3 * n, sizeof(double));

#pragma omp parallel reduction(+ : e_lj, e_c)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *f = force ? fpriv + (size_t)tid * 3 * n : NULL;
#pragma omp for schedule(dynamic, 4)
        for (int c = 0; c < ncell; c++) {
            int cz = c % (use_cells ? nc[2] : 1);
            int cy = use_cells ? (c / nc[2]) % nc[1] : 0;
            int cx = use_cells ? c / (nc[1] * nc[2]) : 0;
            int nneigh = use_cells ? 27 : 1;
            for (int nb = 0; nb < nneigh; nb++) {
                int c2 = 0;
                if (use_cells) {
                    int dx = nb / 9 - 1, dy = (nb / 3) % 3 - 1, dz = nb % 3 - 1;
                    int x2 = (cx + dx + nc[0]) % nc[0], y2 = (cy + dy + nc[1]) % nc[1], z2 = (cz + dz + nc[2]) % nc[2];
                    c2 = (x2 * nc[1] + y2) * nc[2] + z2;
                }
                for (int a = cstart[c]; a < cstart[c + 1]; a++) {
                    int i = corder[a];
                    for (int b = cstart[c2]; b < cstart[c2 + 1]; b++) {
                        int j = corder[b];
                        if (j <= i) continue; /* each unordered pair once (cells are distinct since nc >= 3) */
                        double d0 = min_image(wp[3 * i] - wp[3 * j], L[0]);
                        double d1 = min_image(wp[3 * i + 1] - wp[3 * j + 1], L[1]);
                        double d2 = min_image(wp[3 * i + 2] - wp[3 * j + 2], L[2]);
                        double r2 = d0 * d0 + d1 * d1 + d2 * d2;
                        if (r2 >= rc2) continue;
                        if (is_excluded(&ex, i, j)) continue;
                        double elj, ec, fr;
                        pair_terms(r2, s->charge[i] * s->charge[j], 0.5 * (s->sigma[i] + s->sigma[j]),
                                   sqrt(s->epsilon[i] * s->epsilon[j]), s->ewald_alpha, &elj, &ec, &fr);
                        e_lj += elj;
                        e_c += ec;
                        if (f) {
                            f[3 * i] += fr * d0; f[3 * i + 1] += fr * d1; f[3 * i + 2] += fr * d2;
                            f[3 * j] -= fr * d0; f[3 * j + 1] -= fr * d1; f[3 * j + 2] -= fr * d2;
                        }
                    }
                }
            }
        }
    }
    if (force) {
        for (int t = 0; t < nthreads; t++)
            for (int k = 0; k < 3 * n; k++) force[k] += fpriv[(size_t)t * 3 * n + k];
        free(fpriv);
    }

    /* Ewald exclusion correction: every excluded pair was implicitly included in the reciprocal sum, so
     * -qq*erf(alpha r)/r is added in direct space (minimum-image distance, no cutoff test). */
    double e_x = 0.0;
    for (int k = 0; k < s->n_excl; k++) {
        int i = s->excl[2 * k], j = s->excl[2 * k + 1];
        double d0 = min_image(pos[3 * i] - pos[3 * j], L[0]);
        double d1 = min_image(pos[3 * i + 1] - pos[3 * j + 1], L[1]);
        double d2 = min_image(pos[3 * i + 2] - pos[3 * j + 2], L[2]);
        double r2 = d0 * d0 + d1 * d1 + d2 * d2;
        double r = sqrt(r2);
        double qq = ATM_ORACLE_ONE_4PI_EPS0 * s->charge[i] * s->charge[j];
        if (r == 0.0 || qq == 0.0 || s->ewald_alpha == 0.0) continue;
        double ar = s->ewald_alpha * r;
        double erf_ar = erf(ar);
        e_x -= qq * erf_ar / r;
        if (force) {
            double fr = -qq / r * (erf_ar - ar * exp(-ar * ar) * M_2_SQRTPI) / r2;
            force[3 * i] += fr * d0; force[3 * i + 1] += fr * d1; force[3 * i + 2] += fr * d2;
            force[3 * j] -= fr * d0; force[3 * j + 1] -= fr * d1; force[3 * j + 2] -= fr * d2;
        }
    }

    /* 1-4 exceptions: plain Coulomb with chargeProd and LJ with the exception's sigma/epsilon,
     * no cutoff, no periodic image (OpenMM default). */
    double e_14 = 0.0;
    for (int k = 0; k < s->n_exc14; k++) {
        int i = s->exc14[2 * k], j = s->exc14[2 * k + 1];
        double d0 = pos[3 * i] - pos[3 * j], d1 = pos[3 * i + 1] - pos[3 * j + 1], d2 = pos[3 * i + 2] - pos[3 * j + 2];
        double r2 = d0 * d0 + d1 * d1 + d2 * d2;
        double r = sqrt(r2);
        double qq = ATM_ORACLE_ONE_4PI_EPS0 * s->exc14_par[3 * k];
        double sig = s->exc14_par[3 * k + 1], eps = s->exc14_par[3 * k + 2];
        double s2 = sig * sig / r2, s6 = s2 * s2 * s2;
        e_14 += 4.0 * eps * s6 * (s6 - 1.0) + qq / r;
        if (force) {
            double fr = (24.0 * eps * s6 * (2.0 * s6 - 1.0) + qq / r) / r2;
            force[3 * i] += fr * d0; force[3 * i + 1] += fr * d1; force[3 * i + 2] += fr * d2;
            force[3 * j] -= fr * d0; force[3 * j + 1] -= fr * d1; force[3 * j + 2] -= fr * d2;
        }
    }

    free(ex.start); free(ex.list); free(cell_of); free(cstart); free(corder); free(wp);
    if (out) { out[0] = e_lj; out[1] = e_c; out[2] = e_x; out[3] = e_14; }
    return e_lj + e_c + e_x + e_14;
}

/* Exact Ewald reciprocal energy: E = (2 pi / V) k_e sum_{k != 0} exp(-k^2/4a^2)/k^2 |S(k)|^2,
 * S(k) = sum_j q_j exp(i k.r_j).  Half space (k and -k folded) with weight 2. */
double atm_oracle_ewald_recip(const atm_oracle_system *s, const double *pos, double tol, double *force) {
    const int n = s->n;
    const double *L = s->box;
    const double a = s->ewald_alpha;
    const double V = L[0] * L[1] * L[2];
    const double kcut2 = -4.0 * a * a * log(tol);
    int kmax[3];
    for (int d = 0; d < 3; d++) kmax[d] = (int)ceil(sqrt(kcut2) * L[d] / (2.0 * M_PI));
    const int nx = kmax[0], ny = kmax[1], nz = kmax[2];
    long nk = (long)(nx + 1) * (2 * ny + 1) * (2 * nz + 1);
    double energy = 0.0;
    int nthreads = atm_oracle_num_threads();
    double *fpriv = force ? (double *)calloc((size_t)nthreads * 3 * n, sizeof(double)) : NULL;

#pragma omp parallel reduction(+ : energy)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *f = force ? fpriv + (size_t)tid * 3 * n : NULL;
        double *cs = (double *)malloc(sizeof(double) * 2 * (size_t)n);
#pragma omp for schedule(dynamic, 16)
        for (long t = 0; t < nk; t++) {
            int ix = (int)(t / ((2 * ny + 1) * (2 * nz + 1)));
            int iy = (int)((t / (2 * nz + 1)) % (2 * ny + 1)) - ny;
            int iz = (int)(t % (2 * nz + 1)) - nz;
            /* half space: ix > 0, or ix == 0 and (iy > 0 or (iy == 0 and iz > 0)) */
            if (ix == 0 && (iy < 0 || (iy == 0 && iz <= 0))) continue;
            double kx = 2.0 * M_PI * ix / L[0], ky = 2.0 * M_PI * iy / L[1], kz = 2.0 * M_PI * iz / L[2];
            double k2 = kx * kx + ky * ky + kz * kz;
            if (k2 > kcut2) continue;
            double sr = 0.0, si = 0.0;
            for (int j = 0; j < n; j++) {
                double ph = kx * pos[3 * j] + ky * pos[3 * j + 1] + kz * pos[3 * j + 2];
                double c = cos(ph), sn = sin(ph);
                cs[2 * j] = c; cs[2 * j + 1] = sn;
                sr += s->charge[j] * c;
                si += s->charge[j] * sn;
            }
            double ak = 2.0 * (2.0 * M_PI / V) * ATM_ORACLE_ONE_4PI_EPS0 * exp(-k2 / (4.0 * a * a)) / k2;
            energy += ak * (sr * sr + si * si);
            if (f) {
                for (int j = 0; j < n; j++) {
                    /* -dE/dr_j = 2 ak q_j k (sin_j * Sr - cos_j * Si) */
                    double g = 2.0 * ak * s->charge[j] * (cs[2 * j + 1] * sr - cs[2 * j] * si);
                    f[3 * j] += g * kx; f[3 * j + 1] += g * ky; f[3 * j + 2] += g * kz;
                }
            }
        }
        free(cs);
    }
    if (force) {
        for (int t = 0; t < nthreads; t++)
            for (int k = 0; k < 3 * n; k++) force[k] += fpriv[(size_t)t * 3 * n + k];
        free(fpriv);
    }
    return energy;
}

/* ------------------------------------------------------------------------------------------------
 * Smooth PME reciprocal space (double precision, naive separable DFT -- this is a checker, not a fast code)
 * ---------------------------------------------------------------------------------------------- */

/* cardinal B-spline weights M_p(w + p-1-k), k = 0..p-1, and their derivatives, by the Essmann recursion */
static void bspline(double w, int order, double *theta, double *dtheta) {
    double a[16];
    for (int k = 0; k < order; k++) a[k] = 0.0;
    a[1] = w;
    a[0] = 1.0 - w;
    for (int k = 3; k < order; k++) {
        double div = 1.0 / (k - 1.0);
        a[k - 1] = div * w * a[k - 2];
        for (int j = 1; j <= k - 2; j++) a[k - j - 1] = div * ((w + j) * a[k - j - 2] + (k - j - w) * a[k - j - 1]);
        a[0] = div * (1.0 - w) * a[0];
    }
    /* derivative from the order-1 spline */
    dtheta[0] = -a[0];
    for (int k = 1; k < order; k++) dtheta[k] = a[k - 1] - a[k];
    /* last recursion step */
    {
        double div = 1.0 / (order - 1.0);
        a[order - 1] = div * w * a[order - 2];
        for (int j = 1; j <= order - 2; j++)
            a[order - j - 1] = div * ((w + j) * a[order - j - 2] + (order - j - w) * a[order - j - 1]);
        a[0] = div * (1.0 - w) * a[0];
    }
    for (int k = 0; k < order; k++) theta[k] = a[k];
}

/* in-place separable DFT of a complex grid g[n0][n1][n2] (re, im interleaved); sign = -1 forward, +1 backward (unscaled) */
static void dft3(double *g, const int n[3], int sign) {
    for (int axis = 0; axis < 3; axis++) {
        const int len = n[axis];
        double *cs = (double *)malloc(sizeof(double) * 2 * (size_t)len);
        for (int k = 0; k < len; k++) {
            cs[2 * k] = cos(2.0 * M_PI * k / len);
            cs[2 * k + 1] = sign * sin(2.0 * M_PI * k / len);
        }
        const size_t stride = axis == 0 ? (size_t)n[1] * n[2] : (axis == 1 ? (size_t)n[2] : 1);
        const size_t nlines = (size_t)n[0] * n[1] * n[2] / len;
#pragma omp parallel
        {
            double *line = (double *)malloc(sizeof(double) * 4 * (size_t)len);
#pragma omp for
            for (long li = 0; li < (long)nlines; li++) {
                size_t base;
                if (axis == 0) base = (size_t)li;
                else if (axis == 1) base = ((size_t)li / n[2]) * (size_t)n[1] * n[2] + (size_t)li % n[2];
                else base = (size_t)li * n[2];
                for (int k = 0; k < len; k++) {
                    line[2 * k] = g[2 * (base + k * stride)];
                    line[2 * k + 1] = g[2 * (base + k * stride) + 1];
                }
                double *out = line + 2 * len;
                for (int m = 0; m < len; m++) {
                    double re = 0.0, im = 0.0;
                    for (int k = 0; k < len; k++) {
                        const int t = (int)(((long)m * k) % len);
                        re += line[2 * k] * cs[2 * t] - line[2 * k + 1] * cs[2 * t + 1];
                        im += line[2 * k] * cs[2 * t + 1] + line[2 * k + 1] * cs[2 * t];
                    }
                    out[2 * m] = re;
                    out[2 * m + 1] = im;
                }
                for (int k = 0; k < len; k++) {
                    g[2 * (base + k * stride)] = out[2 * k];
                    g[2 * (base + k * stride) + 1] = out[2 * k + 1];
                }
            }
            free(line);
        }
        free(cs);
    }
}

/* |b(m)|^2 of the Euler exponential spline for one dimension */
static void bspline_moduli(int n, int order, double *mod) {
    double theta[16], dtheta[16];
    bspline(0.0, order, theta, dtheta);  /* theta[k] = M_p(p-1-k) ... values at the integers */
    for (int m = 0; m < n; m++) {
        double sr = 0.0, si = 0.0;
        for (int k = 0; k < order; k++) {
            const double arg = 2.0 * M_PI * m * k / n;
            sr += theta[k] * cos(arg);
            si += theta[k] * sin(arg);
        }
        mod[m] = sr * sr + si * si;
    }
    for (int m = 0; m < n; m++)
        if (mod[m] < 1e-7) mod[m] = 0.5 * (mod[(m + n - 1) % n] + mod[(m + 1) % n]);
}

double atm_oracle_pme_recip(const atm_oracle_system *s, const double *pos, const int n[3], int order, double *force) {
    const int N = s->n;
    const double *L = s->box;
    const double V = L[0] * L[1] * L[2];
    const size_t ng = (size_t)n[0] * n[1] * n[2];
    double *g = (double *)calloc(2 * ng, sizeof(double));
    double *th = (double *)malloc(sizeof(double) * 3 * 16 * (size_t)N), *dth = (double *)malloc(sizeof(double) * 3 * 16 * (size_t)N);
    int *k0 = (int *)malloc(sizeof(int) * 3 * (size_t)N);
    /* spread */
    for (int i = 0; i < N; i++) {
        for (int d = 0; d < 3; d++) {
            double u = pos[3 * i + d] / L[d];
            u = (u - floor(u)) * n[d];
            int fl = (int)floor(u);
            if (fl >= n[d]) fl = n[d] - 1;
            bspline(u - fl, order, th + (3 * (size_t)i + d) * 16, dth + (3 * (size_t)i + d) * 16);
            k0[3 * i + d] = fl - order + 1;
        }
        const double q = s->charge[i];
        for (int a = 0; a < order; a++) {
            const int ia = ((k0[3 * i] + a) % n[0] + n[0]) % n[0];
            for (int b = 0; b < order; b++) {
                const int ib = ((k0[3 * i + 1] + b) % n[1] + n[1]) % n[1];
                const double qab = q * th[(3 * (size_t)i) * 16 + a] * th[(3 * (size_t)i + 1) * 16 + b];
                for (int c = 0; c < order; c++) {
                    const int ic = ((k0[3 * i + 2] + c) % n[2] + n[2]) % n[2];
                    g[2 * (((size_t)ia * n[1] + ib) * n[2] + ic)] += qab * th[(3 * (size_t)i + 2) * 16 + c];
                }
            }
        }
    }
    dft3(g, n, -1);
    /* convolution with exp(-pi^2 m^2 / alpha^2) / (pi V m^2 B(m)) and the energy */
    double *mod[3];
    for (int d = 0; d < 3; d++) {
        mod[d] = (double *)malloc(sizeof(double) * (size_t)n[d]);
        bspline_moduli(n[d], order, mod[d]);
    }
    const double fac = M_PI * M_PI / (s->ewald_alpha * s->ewald_alpha);
    double energy = 0.0;
    for (int a = 0; a < n[0]; a++) {
        const double ma = (a <= n[0] / 2 ? a : a - n[0]) / L[0];
        for (int b = 0; b < n[1]; b++) {
            const double mb = (b <= n[1] / 2 ? b : b - n[1]) / L[1];
            for (int c = 0; c < n[2]; c++) {
                const size_t idx = ((size_t)a * n[1] + b) * n[2] + c;
                if (a == 0 && b == 0 && c == 0) { g[2 * idx] = g[2 * idx + 1] = 0.0; continue; }
                const double mc = (c <= n[2] / 2 ? c : c - n[2]) / L[2];
                const double m2 = ma * ma + mb * mb + mc * mc;
                const double eterm = ATM_ORACLE_ONE_4PI_EPS0 * exp(-fac * m2) / (M_PI * V * m2 * mod[0][a] * mod[1][b] * mod[2][c]);
                energy += 0.5 * eterm * (g[2 * idx] * g[2 * idx] + g[2 * idx + 1] * g[2 * idx + 1]);
                g[2 * idx] *= eterm;
                g[2 * idx + 1] *= eterm;
            }
        }
    }
    if (force) {
        dft3(g, n, +1); /* potential on the grid (real part) */
        for (int i = 0; i < N; i++) {
            const double q = s->charge[i];
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
            for (int a = 0; a < order; a++) {
                const int ia = ((k0[3 * i] + a) % n[0] + n[0]) % n[0];
                const double ta = th[(3 * (size_t)i) * 16 + a], da = dth[(3 * (size_t)i) * 16 + a];
                for (int b = 0; b < order; b++) {
                    const int ib = ((k0[3 * i + 1] + b) % n[1] + n[1]) % n[1];
                    const double tb = th[(3 * (size_t)i + 1) * 16 + b], db = dth[(3 * (size_t)i + 1) * 16 + b];
                    for (int c = 0; c < order; c++) {
                        const int ic = ((k0[3 * i + 2] + c) % n[2] + n[2]) % n[2];
                        const double tc = th[(3 * (size_t)i + 2) * 16 + c], dc = dth[(3 * (size_t)i + 2) * 16 + c];
                        const double phi = g[2 * (((size_t)ia * n[1] + ib) * n[2] + ic)];
                        f0 += da * tb * tc * phi;
                        f1 += ta * db * tc * phi;
                        f2 += ta * tb * dc * phi;
                    }
                }
            }
            force[3 * i] -= q * f0 * n[0] / L[0];
            force[3 * i + 1] -= q * f1 * n[1] / L[1];
            force[3 * i + 2] -= q * f2 * n[2] / L[2];
        }
    }
    for (int d = 0; d < 3; d++) free(mod[d]);
    free(g); free(th); free(dth); free(k0);
    return energy;
}

/* ATMMetaForceImpl.cpp:110-122 on the CPU: copyState, inner evaluation 1, inner evaluation 2, execute. */
void atm_oracle_step(const atm_oracle_system *s, const double p[9], const double *pos, const double *displ,
                     double du_ext, double *force_out, double energies[5]) {
    const int n = s->n;
    double *pos1 = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double *pos2 = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double *f1 = (double *)calloc(3 * (size_t)n, sizeof(double));
    double *f2 = (double *)calloc(3 * (size_t)n, sizeof(double));
    atm_oracle_copy_state_ref(n, pos, displ, pos1, pos2);
    double U1 = atm_oracle_nb_direct(s, pos1, f1, NULL);
    double U2 = atm_oracle_nb_direct(s, pos2, f2, NULL) + du_ext;
    double sc[7];
    atm_oracle_scalars(p, U1, U2, sc);
    if (force_out) atm_oracle_merge_ref(n, force_out, f1, f2, sc[6], p[8]);
    energies[0] = U1; energies[1] = U2; energies[2] = sc[0]; energies[3] = sc[4]; energies[4] = sc[5];
    free(pos1); free(pos2); free(f1); free(f2);
}
