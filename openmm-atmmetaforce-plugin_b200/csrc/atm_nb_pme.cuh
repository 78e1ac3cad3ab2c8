// atm_nb_pme.cuh -- Tier 2, optional: two-state smooth PME reciprocal space (spread, finalize, convolve, gather; the
// transforms are cuFFT calls in atm_nb.cu).  Included by atm_nb.cu only.
#pragma once

#include "atm_common.cuh"
#include "atm_nb_types.cuh"

namespace atm {

// ------------------------------------------------------------------------------------------------
// Two-state smooth PME reciprocal space.  Works on the cluster-order sites in double:
// the environment is spread ONCE; the displaced atoms (state 1) and their ghosts (state 2) are spread into two small
// extra accumulators; Q1 = env + lig, Q2 = env + ghost.  Two batched FFT pairs give the two potentials; environment
// sites gather from both (their reciprocal force differs between the states because the ligand's field moved).
// Charges carry sqrt(k_e), so the influence function needs no Coulomb constant.  Grid accumulation is 2^40 fixed
// point (deterministic), everything after it double precision: U2 - U1 keeps its digits.
// ------------------------------------------------------------------------------------------------
constexpr int PME_MAX_ORDER = 8;

// cardinal B-spline weights theta[k] and derivatives dtheta[k], k = 0..ORDER-1, for fractional offset w (Essmann 1995);
// ORDER is a compile-time constant so that everything stays in registers
template <int ORDER>
__device__ __forceinline__ void pme_bspline(double w, double (&theta)[ORDER], double (&dtheta)[ORDER]) {
#pragma unroll
    for (int k = 0; k < ORDER; k++) theta[k] = 0.0;
    theta[1] = w;
    theta[0] = 1.0 - w;
#pragma unroll
    for (int k = 3; k < ORDER; k++) {
        const double div = 1.0 / (k - 1.0);
        theta[k - 1] = div * w * theta[k - 2];
#pragma unroll
        for (int j = 1; j <= k - 2; j++) theta[k - j - 1] = div * ((w + j) * theta[k - j - 2] + (k - j - w) * theta[k - j - 1]);
        theta[0] = div * (1.0 - w) * theta[0];
    }
    dtheta[0] = -theta[0];
#pragma unroll
    for (int k = 1; k < ORDER; k++) dtheta[k] = theta[k - 1] - theta[k];
    const double div = 1.0 / (ORDER - 1.0);
    theta[ORDER - 1] = div * w * theta[ORDER - 2];
#pragma unroll
    for (int j = 1; j <= ORDER - 2; j++)
        theta[ORDER - j - 1] = div * ((w + j) * theta[ORDER - j - 2] + (ORDER - j - w) * theta[ORDER - j - 1]);
    theta[0] = div * (1.0 - w) * theta[0];
}

template <int ORDER>
struct PmeSite {
    int k0[3];
    double th[3][ORDER], dth[3][ORDER];
};

template <int ORDER>
__device__ __forceinline__ void pme_site_setup(const NbDev &d, const float4 &x, const float4 &L, PmeSite<ORDER> &ps) {
    const int n[3] = {d.gx, d.gy, d.gz};
    const double xr[3] = {(double)x.x / (double)L.x, (double)x.y / (double)L.y, (double)x.z / (double)L.z};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double u = (xr[c] - floor(xr[c])) * n[c];
        int fl = (int)floor(u);
        if (fl >= n[c]) fl = n[c] - 1;
        pme_bspline<ORDER>(u - fl, ps.th[c], ps.dth[c]);
        ps.k0[c] = fl - ORDER + 1;
    }
}

__device__ __forceinline__ int pme_wrap(int i, int n) {
    i += i < 0 ? n : 0;
    return i - (i >= n ? n : 0);
}

// One thread per site slot: ORDER^3 fixed-point atomics.  Two accumulators per replica:
//   acc[0] = Q1 = environment + displaced atoms,  acc[1] = Q2 - Q1 = ghosts - displaced atoms
// so the environment (almost every site) is spread exactly once.
// (A cooperative variant -- B-spline weights staged in shared memory, the block walking the (site, grid point) items
// with z fastest so that a warp-wide RED touches ~13 sectors instead of 32 -- was measured and is SLOWER, 295 vs 257 us
// at 22 replicas: the limit is the L2 atomic-operation rate (64 M 64-bit REDs per launch), not the sector count.)
template <int ORDER>
__global__ void __launch_bounds__(128) pme_spread_kernel(NbDev d) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (s >= CL * d.nclusters[r]) return;
    const size_t rs = (size_t)r * d.Smax + s;
    if (d.slot_site[rs] < 0) return;
    const int cls = d.cmeta[(size_t)r * d.Cmax + (s >> 3)] & 0xffff;
    const int kind = class_kind(cls, d.G);
    const float4 x = d.xs[rs];
    PmeSite<ORDER> ps;
    pme_site_setup<ORDER>(d, x, d.box[r], ps);
    const size_t ng = (size_t)d.gx * d.gy * d.gz;
    unsigned long long *acc1 = d.pme_acc + (size_t)r * 2 * ng, *accd = acc1 + ng;
    const double q = (double)x.w * PME_SCALE;
#pragma unroll
    for (int a = 0; a < ORDER; a++) {
        const int ia = pme_wrap(ps.k0[0] + a, d.gx);
#pragma unroll
        for (int b = 0; b < ORDER; b++) {
            const int ib = pme_wrap(ps.k0[1] + b, d.gy);
            const double qab = q * ps.th[0][a] * ps.th[1][b];
            const size_t row = ((size_t)ia * d.gy + ib) * d.gz;
#pragma unroll
            for (int c = 0; c < ORDER; c++) {
                const int ic = pme_wrap(ps.k0[2] + c, d.gz);
                const long long v = __double2ll_rn(qab * ps.th[2][c]);
                if (kind != 2) atomicAdd(acc1 + row + ic, (unsigned long long)v);        // environment, displaced atoms -> Q1
                if (kind == 1) atomicAdd(accd + row + ic, (unsigned long long)(-v));     // displaced atoms leave in state 2
                if (kind == 2) atomicAdd(accd + row + ic, (unsigned long long)v);        // ghosts arrive in state 2
            }
        }
    }
}

// Q1, Q2 = Q1 + (Q2 - Q1) as doubles; the accumulators are handed back zeroed
__global__ void pme_finalize_kernel(NbDev d) {
    const size_t ng = (size_t)d.gx * d.gy * d.gz;
    const size_t i = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);  // two cells per thread: 128-bit accesses
    const int r = blockIdx.y;
    if (i >= ng) return;
    unsigned long long *acc = d.pme_acc + (size_t)r * 2 * ng;
    double *grid = d.pme_grid + (size_t)r * 2 * ng;
    if (i + 1 < ng && (ng & 1) == 0) {
        const ulonglong2 a1 = *reinterpret_cast<const ulonglong2 *>(acc + i), ad = *reinterpret_cast<const ulonglong2 *>(acc + ng + i);
        *reinterpret_cast<ulonglong2 *>(acc + i) = make_ulonglong2(0ull, 0ull);
        if (ad.x != 0ull || ad.y != 0ull) *reinterpret_cast<ulonglong2 *>(acc + ng + i) = make_ulonglong2(0ull, 0ull);
        const long long q1x = (long long)a1.x, q1y = (long long)a1.y;
        *reinterpret_cast<double2 *>(grid + i) = make_double2((double)q1x * (1.0 / PME_SCALE), (double)q1y * (1.0 / PME_SCALE));
        *reinterpret_cast<double2 *>(grid + ng + i) = make_double2((double)(q1x + (long long)ad.x) * (1.0 / PME_SCALE),
                                                                   (double)(q1y + (long long)ad.y) * (1.0 / PME_SCALE));
    } else {
        for (size_t k = i; k < ng && k < i + 2; k++) {
            const long long q1 = (long long)acc[k], dq = (long long)acc[ng + k];
            acc[k] = 0ull;
            if (dq != 0) acc[ng + k] = 0ull;
            grid[k] = (double)q1 * (1.0 / PME_SCALE);
            grid[ng + k] = (double)(q1 + dq) * (1.0 / PME_SCALE);
        }
    }
}

// multiply the spectra by exp(-pi^2 m^2/alpha^2) / (pi V m^2 B(m)); accumulate the two reciprocal energies
__global__ void __launch_bounds__(256) pme_convolve_kernel(NbDev d) {
    const int nzh = d.gz / 2 + 1;
    const size_t nspec = (size_t)d.gx * d.gy * nzh;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y, state = blockIdx.z;
    double en = 0.0;
    if (i < nspec) {
        const int c = (int)(i % nzh), b = (int)((i / nzh) % d.gy), a = (int)(i / ((size_t)nzh * d.gy));
        double2 *spec = d.pme_spec + ((size_t)r * 2 + state) * nspec;
        if (a == 0 && b == 0 && c == 0) {
            spec[i] = make_double2(0.0, 0.0);
        } else {
            const float4 L = d.box[r];
            const double ma = (double)(a <= d.gx / 2 ? a : a - d.gx) / (double)L.x, mb = (double)(b <= d.gy / 2 ? b : b - d.gy) / (double)L.y,
                         mc = (double)c / (double)L.z;
            const double m2 = ma * ma + mb * mb + mc * mc;
            const double V = (double)L.x * (double)L.y * (double)L.z;
            const double fac = 9.869604401089358 / ((double)d.alpha * (double)d.alpha);  // pi^2 / alpha^2
            const double eterm = exp(-fac * m2) / (3.141592653589793 * V * m2 * d.pme_mod[a] * d.pme_mod[d.gx + b] * d.pme_mod[d.gx + d.gy + c]);
            double2 v = spec[i];
            const double w = (c == 0 || (2 * c == d.gz)) ? 1.0 : 2.0;  // half spectrum: the conjugate half counts too
            en = 0.5 * w * eterm * (v.x * v.x + v.y * v.y);
            v.x *= eterm; v.y *= eterm;
            spec[i] = v;
        }
    }
    // block reduction, one fixed-point atomic per block
    __shared__ double red[256 / 32];
    for (int off = 16; off > 0; off >>= 1) en += __shfl_xor_sync(0xffffffffu, en, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = en;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 256 / 32; k++) t += red[k];
        atomicAdd(d.eacc + (size_t)r * EACC_SLOTS + 6 + state, (unsigned long long)__double2ll_rn(t * ENERGY_SCALE));
    }
}

// one thread per site: F = -q (n/L) sum dtheta theta theta phi, into the state-specific accumulators
template <int ORDER>
__global__ void __launch_bounds__(128, 4) pme_gather_kernel(NbDev d) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (s >= CL * d.nclusters[r]) return;
    const size_t rs = (size_t)r * d.Smax + s;
    if (d.slot_site[rs] < 0) return;
    const int cls = d.cmeta[(size_t)r * d.Cmax + (s >> 3)] & 0xffff;
    const int kind = class_kind(cls, d.G);
    const float4 x = d.xs[rs];
    const float4 L = d.box[r];
    PmeSite<ORDER> ps;
    pme_site_setup<ORDER>(d, x, L, ps);
    const size_t ng = (size_t)d.gx * d.gy * d.gz;
    const double *phi1 = d.pme_grid + (size_t)r * 2 * ng, *phi2 = phi1 + ng;
    double f1x = 0, f1y = 0, f1z = 0, f2x = 0, f2y = 0, f2z = 0;
    const bool want1 = kind != 2, want2 = kind != 1;
#pragma unroll
    for (int a = 0; a < ORDER; a++) {
        const int ia = pme_wrap(ps.k0[0] + a, d.gx);
#pragma unroll
        for (int b = 0; b < ORDER; b++) {
            const int ib = pme_wrap(ps.k0[1] + b, d.gy);
            const size_t row = ((size_t)ia * d.gy + ib) * d.gz;
            const double tx = ps.dth[0][a] * ps.th[1][b], ty = ps.th[0][a] * ps.dth[1][b], tz = ps.th[0][a] * ps.th[1][b];
#pragma unroll
            for (int c = 0; c < ORDER; c++) {
                const int ic = pme_wrap(ps.k0[2] + c, d.gz);
                const double wx = tx * ps.th[2][c], wy = ty * ps.th[2][c], wz = tz * ps.dth[2][c];
                if (want1) { const double p = __ldg(phi1 + row + ic); f1x += wx * p; f1y += wy * p; f1z += wz * p; }
                if (want2) { const double p = __ldg(phi2 + row + ic); f2x += wx * p; f2y += wy * p; f2z += wz * p; }
            }
        }
    }
    const double q = (double)x.w;
    const double sx = -q * d.gx / (double)L.x, sy = -q * d.gy / (double)L.y, sz = -q * d.gz / (double)L.z;
    const size_t cs = (size_t)d.R * d.Smax, rsite = (size_t)r * d.Smax;
    unsigned long long *buf1 = d.buf + 3 * cs + rsite, *buf2 = d.buf + 6 * cs + rsite;
    if (want1) {
        atomicAdd(buf1 + s, (unsigned long long)__double2ll_rn(sx * f1x * FORCE_SCALE));
        atomicAdd(buf1 + cs + s, (unsigned long long)__double2ll_rn(sy * f1y * FORCE_SCALE));
        atomicAdd(buf1 + 2 * cs + s, (unsigned long long)__double2ll_rn(sz * f1z * FORCE_SCALE));
    }
    if (want2) {
        atomicAdd(buf2 + s, (unsigned long long)__double2ll_rn(sx * f2x * FORCE_SCALE));
        atomicAdd(buf2 + cs + s, (unsigned long long)__double2ll_rn(sy * f2y * FORCE_SCALE));
        atomicAdd(buf2 + 2 * cs + s, (unsigned long long)__double2ll_rn(sz * f2z * FORCE_SCALE));
    }
}

}  // namespace atm
