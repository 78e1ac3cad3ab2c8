// atm_nb_pme.cuh -- Tier 2, optional: two-state smooth PME reciprocal space (spread, finalize, convolve, gather; the
// transforms are cuFFT calls in atm_nb.cu).  Included by atm_nb.cu only.
#pragma once

#include "atm_common.cuh"
#include "atm_nb_types.cuh"
#include "atm_nb_force.cuh"   // scalar_stage_replica (the blend kernel forms sp before the inverse transform)

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
template <int ORDER, typename T>
__device__ __forceinline__ void pme_bspline(T w, T (&theta)[ORDER], T (&dtheta)[ORDER]) {
#pragma unroll
    for (int k = 0; k < ORDER; k++) theta[k] = T(0);
    theta[1] = w;
    theta[0] = T(1) - w;
#pragma unroll
    for (int k = 3; k < ORDER; k++) {
        const T div = T(1) / T(k - 1);
        theta[k - 1] = div * w * theta[k - 2];
#pragma unroll
        for (int j = 1; j <= k - 2; j++) theta[k - j - 1] = div * ((w + T(j)) * theta[k - j - 2] + (T(k - j) - w) * theta[k - j - 1]);
        theta[0] = div * (T(1) - w) * theta[0];
    }
    dtheta[0] = -theta[0];
#pragma unroll
    for (int k = 1; k < ORDER; k++) dtheta[k] = theta[k - 1] - theta[k];
    const T div = T(1) / T(ORDER - 1);
    theta[ORDER - 1] = div * w * theta[ORDER - 2];
#pragma unroll
    for (int j = 1; j <= ORDER - 2; j++)
        theta[ORDER - j - 1] = div * ((w + T(j)) * theta[ORDER - j - 2] + (T(ORDER - j) - w) * theta[ORDER - j - 1]);
    theta[0] = div * (T(1) - w) * theta[0];
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
        pme_bspline<ORDER, double>(u - fl, ps.th[c], ps.dth[c]);
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

// ================================================================================================
// Single-precision mesh pipeline (default; the double kernels above stay selectable with ATM_B200_PME_F64=1):
//   pme_spread_tile_kernel   one block OWNS a brick of mesh cells (a tile of xy cells, all z): it scans the sites of the
//                            xy columns that can reach the brick (sites are stored column by column), accumulates their
//                            contributions in a shared-memory tile with 32-bit fixed-point atomics (deterministic), and
//                            stores the brick as floats straight into the transform input -- no global atomics, no
//                            accumulator to clear, no separate conversion pass.  Two passes per block: Q1 (environment +
//                            displaced atoms) and dQ = Q2 - Q1 (ghosts - displaced atoms).
//   cuFFT R2C (float), batched over both meshes of every replica
//   pme_convolve_f_kernel    influence function in double per mode; E1 = 1/2 sum G |Q1^|^2 and, by linearity of the
//                            transform, E2 - E1 = sum G (Re(conj(Q1^) dQ^) + 1/2 |dQ^|^2), both accumulated in double:
//                            the rounding noise of the big mesh never enters the DIFFERENCE as a difference of two sums
//   cuFFT C2R (float): phi1 and dphi = phi2 - phi1
//   pme_blend_kernel         sp is known by now (scalar stage on the complete energies): phi_b = phi1 + sp dphi, written in
//                            a padded row layout in which every z support is a run inside aligned float4s
//   pme_gather_f_kernel      environment sites: ONE mesh (phi_b), force into the common accumulator; displaced atoms /
//                            ghosts: phi1 / phi1 + dphi into the state accumulators.  Float weights, row sums over z first
// Positions -> mesh coordinates stay in double (the float coordinate is the exact input; its fractional mesh offset
// would lose five digits in float); weights, meshes and transforms are float.
// ================================================================================================
constexpr float PME_TILE_SCALE = 16777216.0f;   // 2^24 fixed point of the shared-memory tile (|cell charge| < 128 sqrt(k_e) e)
constexpr int PME_SPREAD_THREADS = 256;

// mesh coordinate of one component: cell index in [0, n) and the fractional offset inside the cell
__device__ __forceinline__ int pme_cell(float x, double invL, int n, double &w) {
    const double xr = (double)x * invL;
    const double u = (xr - floor(xr)) * n;
    int f = (int)u;
    if (f >= n) f = n - 1;
    w = u - f;
    return f;
}

template <int ORDER, bool WANT_D>
__device__ __forceinline__ void pme_site_setup_f(const NbDev &d, const float4 &x, const double (&invL)[3], int (&fl)[3], float (&th)[3][ORDER],
                                                 float (&dth)[3][ORDER]) {
    const int n[3] = {d.gx, d.gy, d.gz};
    const float xc[3] = {x.x, x.y, x.z};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double w;
        fl[c] = pme_cell(xc[c], invL[c], n[c], w);
        pme_bspline<ORDER, float>((float)w, th[c], dth[c]);
    }
}

__device__ __forceinline__ int pme_mod(int i, int n) {   // any i >= -n
    i %= n;
    return i < 0 ? i + n : i;
}

// Does the spline support of mesh coordinate (flx, fly) touch the brick [x0, x0 + tw) x [y0, y0 + th)?
template <int ORDER>
__device__ __forceinline__ bool pme_touches(const NbDev &d, int flx, int fly, int x0, int tw, int y0, int th_) {
    // cells flx - (ORDER - 1) .. flx  <=>  the brick-relative index of the LAST cell is below tw + ORDER - 1
    return pme_wrap(flx - x0, d.gx) < tw + ORDER - 1 && pme_wrap(fly - y0, d.gy) < th_ + ORDER - 1;
}

// Tile layout: one row per (x, y) column of the brick, `zpad` cells (a multiple of 4, >= ORDER - 1) in front of the gz cells
// of the column for supports that start below z = 0 (folded onto the top of the column when the brick is written out), so
// that the z cells of a site are always consecutive; the row stride is a multiple of 4 cells (128-bit write-out) with an odd
// number of 16-byte units (rows of equal z fall into different banks).
__host__ __device__ inline int pme_tile_zpad(int order) { return order <= 5 ? 4 : 8; }
__host__ __device__ inline int pme_tile_stride(int gz, int order) {
    int st = (gz + pme_tile_zpad(order) + 3) & ~3;
    return (st & 4) ? st : st + 4;
}

template <int ORDER>
__device__ __forceinline__ void pme_spread_site(const NbDev &d, const float4 &x, const double (&invL)[3], float sign, int x0, int tw, int y0,
                                                int th_, int stride, int *s_tile) {
    int fl[3];
    float th[3][ORDER], dth[3][ORDER];
    pme_site_setup_f<ORDER, false>(d, x, invL, fl, th, dth);
    // brick-relative index of the first spline cell in x and y (the following ones are consecutive modulo the mesh extent)
    const int tx0 = pme_wrap(pme_wrap(fl[0] - (ORDER - 1), d.gx) - x0, d.gx), ty0 = pme_wrap(pme_wrap(fl[1] - (ORDER - 1), d.gy) - y0, d.gy);
    int yo[ORDER];   // element offset of each y row inside the tile (-1: the row belongs to another block's brick)
    float qz[ORDER];
    const float q = sign * x.w * PME_TILE_SCALE;
#pragma unroll
    for (int c = 0; c < ORDER; c++) {
        qz[c] = q * th[2][c];
        int rb = ty0 + c;
        rb -= rb >= d.gy ? d.gy : 0;
        yo[c] = rb < th_ ? rb * stride : -1;
    }
    int *base = s_tile + (pme_tile_zpad(ORDER) - (ORDER - 1)) + fl[2];   // cell fl - (ORDER - 1) + c of the column lives at zpad + that
#pragma unroll
    for (int a = 0; a < ORDER; a++) {
        int ra = tx0 + a;
        ra -= ra >= d.gx ? d.gx : 0;
        if (ra >= tw) continue;
        int *rowa = base + ra * th_ * stride;
#pragma unroll
        for (int b = 0; b < ORDER; b++) {
            if (yo[b] < 0) continue;
            const float wab = th[0][a] * th[1][b];
            int *row = rowa + yo[b];
#pragma unroll
            for (int c = 0; c < ORDER; c++) atomicAdd(row + c, __float2int_rn(wab * qz[c]));
        }
    }
}

constexpr int PME_SCAN_ROUND = 4 * PME_SPREAD_THREADS;   // candidate slots examined between two looks at the list
constexpr int PME_LIST_CAP = 3 * PME_SCAN_ROUND;          // contributing slots collected before they are spread
constexpr int PME_MAX_RUNS = 256;                         // runs of candidate slots tabulated at a time

template <int ORDER>
__global__ void __launch_bounds__(PME_SPREAD_THREADS, 4) pme_spread_tile_kernel(NbDev d) {
    extern __shared__ int s_tile[];   // [tw][th][stride], then the list of contributing slots (bit 31: negative sign)
    __shared__ int s_count;
    __shared__ int s_run_begin[PME_MAX_RUNS], s_run_end[PME_MAX_RUNS];   // end | bit 31: negative sign
    const int r = blockIdx.y;
    const int tx = blockIdx.x / d.pme_nty, ty = blockIdx.x - tx * d.pme_nty;
    const int x0 = (int)((long long)tx * d.gx / d.pme_ntx), x1 = (int)((long long)(tx + 1) * d.gx / d.pme_ntx);
    const int y0 = (int)((long long)ty * d.gy / d.pme_nty), y1 = (int)((long long)(ty + 1) * d.gy / d.pme_nty);
    const int tw = x1 - x0, th_ = y1 - y0, gz = d.gz;
    const int stride = pme_tile_stride(gz, ORDER), zpad = pme_tile_zpad(ORDER);
    const int ncell = tw * th_ * stride;
    unsigned int *s_list = reinterpret_cast<unsigned int *>(s_tile + d.pme_tile_cells);
    for (int i = threadIdx.x; i < ncell; i += PME_SPREAD_THREADS) s_tile[i] = 0;
    if (threadIdx.x == 0) s_count = 0;
    const float4 L = d.box[r];
    const double invL[3] = {1.0 / (double)L.x, 1.0 / (double)L.y, 1.0 / (double)L.z};
    // xy columns whose sites can reach the brick: a site at mesh coordinate u touches cells floor(u) - (ORDER - 1) .. floor(u),
    // and sits within `margin` of the column it was sorted into at the last rebuild (checked by the gather kernel)
    const float hx = L.x / d.gx, hy = L.y / d.gy, wx = L.x / d.nx, wy = L.y / d.ny;
    const float margin = 0.5f * (d.rlist_outer - sqrtf(d.cutoff2));
    const int cx_lo = (int)floorf((x0 * hx - margin - hx) / wx), cx_hi = (int)floorf(((x1 + ORDER - 1) * hx + margin + hx) / wx);
    const int cy_lo = (int)floorf((y0 * hy - margin - hy) / wy), cy_hi = (int)floorf(((y1 + ORDER - 1) * hy + margin + hy) / wy);
    const int ncx = min(cx_hi - cx_lo + 1, d.nx), ncy = min(cy_hi - cy_lo + 1, d.ny);
    const int cya = pme_mod(cy_lo, d.ny), seg0 = min(ncy, d.ny - cya);   // the y range: [cya, cya + seg0) and, wrapped, [0, ncy - seg0)
    const size_t ng = (size_t)d.gx * d.gy * gz;
    float *out = d.pme_gridf + (size_t)r * 2 * ng;
    const int *bcs = d.bin_cluster_start + (size_t)r * (d.nbins + 1);
    const float4 *xs = d.xs + (size_t)r * d.Smax;
    const int *slot_site = d.slot_site + (size_t)r * d.Smax;
    __syncthreads();
    for (int pass = 0; pass < 2; pass++) {
        // pass 0: Q1 = environment (class 0) + displaced atoms (classes 1..G); pass 1: dQ = ghosts (G+1..2G) - displaced atoms.
        // Runs of candidate slots: one per class, x column and contiguous part of the y range (consecutive bins hold
        // consecutive clusters).  They are tabulated by the threads, then walked by the whole block: slots whose support
        // touches the brick go into the list, and the list is spread -- every lane busy, integer accumulation, so its
        // order does not matter -- before it can overflow and at the end.
        const int cls_lo = pass == 0 ? 0 : 1, ncls = pass == 0 ? d.G + 1 : 2 * d.G;
        const int nruns = ncls * ncx * 2;
        int pending = 0;
        for (int run0 = 0; run0 < nruns; run0 += PME_MAX_RUNS) {
            const int nr = min(PME_MAX_RUNS, nruns - run0);
            if (threadIdx.x < nr) {
                const int t = run0 + threadIdx.x, half = t & 1, ci = t >> 1;
                const int icx = ci % ncx, cls = cls_lo + ci / ncx;
                const int b0 = cls * d.ncol + pme_mod(cx_lo + icx, d.nx) * d.ny + (half ? 0 : cya);
                const int nb = half ? ncy - seg0 : seg0;
                s_run_begin[threadIdx.x] = CL * bcs[b0];
                s_run_end[threadIdx.x] = (CL * bcs[b0 + nb]) | ((pass == 1 && cls <= d.G) ? 0x80000000 : 0);
            }
            __syncthreads();
            for (int k = 0; k < nr; k++) {
                const int begin = s_run_begin[k], e = s_run_end[k];
                const int end = e & 0x7fffffff;
                const unsigned int neg = (unsigned int)e & 0x80000000u;
                for (int base = begin; base < end; base += PME_SCAN_ROUND) {
                    const int len = min(end - base, PME_SCAN_ROUND);
                    if (pending + len > PME_SCAN_ROUND) {   // uniform: look at the list before it can overflow
                        __syncthreads();
                        const int n = s_count;
                        if (n > PME_LIST_CAP - PME_SCAN_ROUND) {
                            for (int i = threadIdx.x; i < n; i += PME_SPREAD_THREADS) {
                                const unsigned int en = s_list[i];
                                pme_spread_site<ORDER>(d, xs[en & 0x7fffffffu], invL, (en >> 31) ? -1.f : 1.f, x0, tw, y0, th_, stride, s_tile);
                            }
                            __syncthreads();
                            if (threadIdx.x == 0) s_count = 0;
                        }
                        __syncthreads();
                        pending = 0;
                    }
                    for (int s = base + threadIdx.x; s < base + len; s += PME_SPREAD_THREADS) {
                        if (slot_site[s] < 0) continue;
                        const float4 x = xs[s];
                        double w;
                        const int flx = pme_cell(x.x, invL[0], d.gx, w), fly = pme_cell(x.y, invL[1], d.gy, w);
                        if (pme_touches<ORDER>(d, flx, fly, x0, tw, y0, th_)) s_list[atomicAdd(&s_count, 1)] = (unsigned int)s | neg;
                    }
                    pending += len;
                }
            }
            __syncthreads();   // the table is rewritten by the next chunk of runs
        }
        {
            const int n = s_count;
            for (int i = threadIdx.x; i < n; i += PME_SPREAD_THREADS) {
                const unsigned int en = s_list[i];
                pme_spread_site<ORDER>(d, xs[en & 0x7fffffffu], invL, (en >> 31) ? -1.f : 1.f, x0, tw, y0, th_, stride, s_tile);
            }
            __syncthreads();
            if (threadIdx.x == 0) s_count = 0;
        }
        // write the brick out and clear the tile for the next pass
        float *o = out + (size_t)pass * ng + ((size_t)x0 * d.gy + y0) * gz;
        if ((gz & 3) == 0) {   // 128-bit: thread per 4 cells
            const int nv = gz >> 2, foldv = (gz - zpad) >> 2, nrow = tw * th_;
            for (int i = threadIdx.x; i < nrow * nv; i += PME_SPREAD_THREADS) {
                const int row = i / nv, v = i - row * nv;
                const int ra = row / th_, rb = row - ra * th_;
                int4 *trow = reinterpret_cast<int4 *>(s_tile + row * stride);
                int4 t = trow[(zpad >> 2) + v];
                if (v >= foldv) {   // supports that started below z = 0
                    const int4 u = trow[v - foldv];
                    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
                }
                const float sc = 1.0f / PME_TILE_SCALE;
                reinterpret_cast<float4 *>(o + ((size_t)ra * d.gy + rb) * gz)[v] = make_float4((float)t.x * sc, (float)t.y * sc, (float)t.z * sc, (float)t.w * sc);
            }
        } else {
            for (int i = threadIdx.x; i < tw * th_ * gz; i += PME_SPREAD_THREADS) {
                const int row = i / gz, iz = i - row * gz;
                const int ra = row / th_, rb = row - ra * th_;
                const int *trow = s_tile + row * stride;
                int v = trow[zpad + iz];
                if (iz >= gz - (ORDER - 1)) v += trow[zpad + iz - gz];
                o[((size_t)ra * d.gy + rb) * gz + iz] = (float)v * (1.0f / PME_TILE_SCALE);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < ncell >> 2; i += PME_SPREAD_THREADS) reinterpret_cast<int4 *>(s_tile)[i] = make_int4(0, 0, 0, 0);
        __syncthreads();
    }
}

// Influence function on the float spectra of Q1 and dQ, and the two reciprocal energies in double.  A warp per four (a, b)
// rows of the half spectrum, lanes over c: G(m) = Ta(a) Tb(b) Tc(c) / (pi V m^2) with T(k) = exp(-pi^2 m_k^2 / alpha^2) / |b(k)|^2;
// the row factors and the z factors are formed once per block in shared memory (one exp per row / per c, not per mode).
constexpr int PME_CONV_WARPS = 8, PME_CONV_RPW = 4, PME_CONV_ROWS = PME_CONV_WARPS * PME_CONV_RPW;   // rows per block
__global__ void __launch_bounds__(32 * PME_CONV_WARPS, 4) pme_convolve_f_kernel(NbDev d) {
    extern __shared__ double s_tc[];   // [nzh] factors along z
    __shared__ double s_tab[PME_CONV_ROWS], s_m2[PME_CONV_ROWS], red[2 * PME_CONV_WARPS];
    const int nzh = d.gz / 2 + 1, nrows = d.gx * d.gy;
    const int r = blockIdx.y, lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const float4 L = d.box[r];
    const double fac = 9.869604401089358 / ((double)d.alpha * (double)d.alpha);  // pi^2 / alpha^2
    // warp 0 forms the row factors while the other warps form the z factors
    if (threadIdx.x >= PME_CONV_ROWS) {
        for (int c = (int)threadIdx.x - PME_CONV_ROWS; c < nzh; c += 32 * PME_CONV_WARPS - PME_CONV_ROWS) {
            const double mc = (double)c / (double)L.z;
            s_tc[c] = exp(-fac * mc * mc) / d.pme_mod[d.gx + d.gy + c];
        }
    } else {
        const int row = blockIdx.x * PME_CONV_ROWS + threadIdx.x;   // a * gy + b
        if (row < nrows) {
            const int a = row / d.gy, b = row - a * d.gy;
            const double ma = (double)(a <= d.gx / 2 ? a : a - d.gx) / (double)L.x, mb = (double)(b <= d.gy / 2 ? b : b - d.gy) / (double)L.y;
            const double mab2 = ma * ma + mb * mb;
            s_m2[threadIdx.x] = mab2;
            s_tab[threadIdx.x] = exp(-fac * mab2) / (3.141592653589793 * (double)L.x * (double)L.y * (double)L.z * d.pme_mod[a] * d.pme_mod[d.gx + b]);
        }
    }
    __syncthreads();
    const size_t nspec = (size_t)nrows * nzh;
    float2 *spec = d.pme_specf + (size_t)r * 2 * nspec;
    const int row0 = blockIdx.x * PME_CONV_ROWS + wrp * PME_CONV_RPW;
    double e1 = 0.0, de = 0.0;
    for (int c0 = 0; c0 < nzh; c0 += 32) {
        const int c = c0 + lane;
        if (c >= nzh) break;
        // the RPW rows of this warp: all loads first, then the arithmetic
        float2 v1[PME_CONV_RPW], vd[PME_CONV_RPW];
#pragma unroll
        for (int k = 0; k < PME_CONV_RPW; k++)
            if (row0 + k < nrows) {
                v1[k] = spec[(size_t)(row0 + k) * nzh + c];
                vd[k] = spec[nspec + (size_t)(row0 + k) * nzh + c];
            }
        const double mc = (double)c / (double)L.z;
        const double tc = s_tc[c], mc2 = mc * mc;
        const double w = (c == 0 || (2 * c == d.gz)) ? 1.0 : 2.0;  // half spectrum: the conjugate half counts too
#pragma unroll
        for (int k = 0; k < PME_CONV_RPW; k++) {
            const int row = row0 + k;
            if (row >= nrows) break;
            float2 *s1 = spec + (size_t)row * nzh + c, *sd = s1 + nspec;
            if (row == 0 && c == 0) {
                *s1 = make_float2(0.f, 0.f);
                *sd = make_float2(0.f, 0.f);
                continue;
            }
            const double eterm = s_tab[wrp * PME_CONV_RPW + k] * tc / (s_m2[wrp * PME_CONV_RPW + k] + mc2);
            const double x1 = v1[k].x, y1 = v1[k].y, xd = vd[k].x, yd = vd[k].y;
            e1 += 0.5 * w * eterm * (x1 * x1 + y1 * y1);
            de += w * eterm * (x1 * xd + y1 * yd + 0.5 * (xd * xd + yd * yd));   // E2 - E1 by linearity of the transform
            const float g = (float)eterm;
            *s1 = make_float2(v1[k].x * g, v1[k].y * g);
            *sd = make_float2(vd[k].x * g, vd[k].y * g);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        e1 += __shfl_xor_sync(0xffffffffu, e1, off);
        de += __shfl_xor_sync(0xffffffffu, de, off);
    }
    if (lane == 0) { red[wrp] = e1; red[PME_CONV_WARPS + wrp] = de; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t1 = 0.0, td = 0.0;
        for (int k = 0; k < PME_CONV_WARPS; k++) { t1 += red[k]; td += red[PME_CONV_WARPS + k]; }
        const long long f1 = __double2ll_rn(t1 * ENERGY_SCALE), fd = __double2ll_rn(td * ENERGY_SCALE);
        atomicAdd(d.eacc + (size_t)r * EACC_SLOTS + 6, (unsigned long long)f1);
        atomicAdd(d.eacc + (size_t)r * EACC_SLOTS + 7, (unsigned long long)(f1 + fd));   // E2 = E1 + (E2 - E1), exactly
    }
}

// The merged force is F1 + sp (F2 - F1) and both are linear in the potentials, so the environment -- almost every site --
// needs only the blended potential phi_b = phi1 + sp dphi: once the reciprocal energies are known (after the convolve
// kernel; the direct-space energies are complete since nb2) the scalar stage gives sp, this kernel writes the blended
// mesh, and the gather reads ONE mesh per environment site instead of two.  The blended mesh has its own row layout,
// made for the gather: every (x, y) row is `zpad` cells (copies of the top of the column) followed by the gz cells of the
// column, in a stride that is a multiple of 4 floats (zero filled), so that the z support of ANY site is a run of
// consecutive floats inside two or three aligned float4s.  The few displaced atoms / ghosts read phi1 and dphi themselves.
// (an odd number of 16-byte units, like the spread tile: rows that start 256 bytes apart would share their L1 sets)
__host__ __device__ inline int pme_blend_stride(int gz, int order) { return pme_tile_stride(gz, order); }

__global__ void __launch_bounds__(256) pme_blend_kernel(NbDev d, const double *__restrict__ energy_ext, int include_energy) {
    __shared__ float s_sp;
    const int r = blockIdx.y;
    if (threadIdx.x == 0) s_sp = (float)scalar_stage_replica(d, r, energy_ext, include_energy, false);   // same inputs as the merge kernel's
    __syncthreads();
    const float sp = s_sp;
    const int gz = d.gz, zpad = pme_tile_zpad(d.pme_order), st = pme_blend_stride(gz, d.pme_order), nv = st >> 2;
    const size_t ng = (size_t)d.gx * d.gy * gz;
    const float *p1 = d.pme_gridf + (size_t)r * 2 * ng, *pd = p1 + ng;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of a padded row
    if (i >= d.gx * d.gy * nv) return;
    const int row = i / nv, v = i - row * nv;
    const float *q1 = p1 + (size_t)row * gz, *qd = pd + (size_t)row * gz;
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int z = 4 * v + k - zpad;   // padded element 4 v + k holds cell z (z < 0: cell gz + z)
        const int zz = z < 0 ? z + gz : z;
        o[k] = z < gz ? fmaf(sp, __ldg(qd + zz), __ldg(q1 + zz)) : 0.f;
    }
    reinterpret_cast<float4 *>(d.pme_blend + ((size_t)r * d.gx * d.gy + row) * st)[v] = make_float4(o[0], o[1], o[2], o[3]);
}

// A displaced atom (kind 1: F1 from phi1) or a ghost (kind 2: F2 from phi1 + dphi): plain loops (a few dozen sites).
template <int ORDER>
__device__ __forceinline__ void pme_gather_special(const NbDev &d, int r, int s, const float4 &x, const float4 &L, int kind) {
    const double invL[3] = {1.0 / (double)L.x, 1.0 / (double)L.y, 1.0 / (double)L.z};
    int fl[3];
    float th[3][ORDER], dth[3][ORDER];   // indexed by loop counters below: these live in local memory, on this rare path only
    pme_site_setup_f<ORDER, true>(d, x, invL, fl, th, dth);
    const size_t ng = (size_t)d.gx * d.gy * d.gz;
    const float *phi1 = d.pme_gridf + (size_t)r * 2 * ng, *phid = phi1 + ng;
    const float wd = kind == 1 ? 0.f : 1.f;
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll 1
    for (int a = 0; a < ORDER; a++) {
        const int ia = pme_wrap(fl[0] - (ORDER - 1) + a, d.gx);
#pragma unroll 1
        for (int b = 0; b < ORDER; b++) {
            const int ib = pme_wrap(fl[1] - (ORDER - 1) + b, d.gy);
            const size_t row = ((size_t)ia * d.gy + ib) * d.gz;
            float s0 = 0.f, sz = 0.f;
#pragma unroll 1
            for (int c = 0; c < ORDER; c++) {
                const int ic = pme_wrap(fl[2] - (ORDER - 1) + c, d.gz);
                const float p = fmaf(wd, __ldg(phid + row + ic), __ldg(phi1 + row + ic));
                s0 = fmaf(th[2][c], p, s0);
                sz = fmaf(dth[2][c], p, sz);
            }
            fx = fmaf(dth[0][a] * th[1][b], s0, fx);
            fy = fmaf(th[0][a] * dth[1][b], s0, fy);
            fz = fmaf(th[0][a] * th[1][b], sz, fz);
        }
    }
    const double q = (double)x.w;
    const size_t cs = (size_t)d.R * d.Smax;
    unsigned long long *buf = d.buf + (kind == 1 ? 3 : 6) * cs + (size_t)r * d.Smax;   // S1 / S2 accumulators
    buf[s] += (unsigned long long)__double2ll_rn(-q * d.gx / (double)L.x * FORCE_SCALE * (double)fx);
    buf[cs + s] += (unsigned long long)__double2ll_rn(-q * d.gy / (double)L.y * FORCE_SCALE * (double)fy);
    buf[2 * cs + s] += (unsigned long long)__double2ll_rn(-q * d.gz / (double)L.z * FORCE_SCALE * (double)fz);
}

// one thread per site.  Environment sites read the blended potential only and add their force to the COMMON accumulator
// (the merge adds it with weight one): per (x, y) row NV aligned float4s that contain the z support, the z weights
// shifted onto that window (zeros outside), row sums first (2 FMA per loaded float), then the xy weights.
#ifndef ATM_PME_GATHER_MINB
#define ATM_PME_GATHER_MINB 6
#endif
template <int ORDER>
__global__ void __launch_bounds__(128, ATM_PME_GATHER_MINB) pme_gather_f_kernel(NbDev d) {
    constexpr int NV = ORDER <= 5 ? 2 : 3;   // float4s that cover ORDER consecutive floats at any 4-byte alignment
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (s >= CL * d.nclusters[r]) return;
    const size_t rs = (size_t)r * d.Smax + s;
    const size_t rc = (size_t)r * d.Cmax + (s >> 3);
    const size_t cs = (size_t)d.R * d.Smax;
    unsigned long long *bufC = d.buf + (size_t)r * d.Smax;
    // every load that does not depend on another one first
    const int site = d.slot_site[rs];
    const int cm = d.cmeta[rc];
    const float4 x = d.xs[rs];
    const float4 cc = d.cc[rc], ch = d.ch[rc];
    const float4 L = d.box[r], iL = d.invbox[r];
    const unsigned long long c0 = bufC[s], c1 = bufC[cs + s], c2 = bufC[2 * cs + s];   // this thread is the only writer of these at this point of the step
    if (site < 0) return;
    const int kind = class_kind(cm & 0xffff, d.G);
    {   // the spread kernel found this site through the xy column it was sorted into at the last rebuild: that holds while
        // no site has left its cluster's bounding box of that time by more than half the outer skin -- the same movement
        // the pair list tolerates.  Beyond it the step is poisoned like a list overflow (flags bit 3).
        const float margin = 0.5f * (d.rlist_outer - sqrtf(d.cutoff2)) + 1e-4f;   // + rounding of the box arithmetic
        if (fabsf(wrap_delta(x.x - cc.x, L.x, iL.x)) - ch.x > margin || fabsf(wrap_delta(x.y - cc.y, L.y, iL.y)) - ch.y > margin)
            atomicOr(&d.flags[0], 8);
    }
    if (kind != 0) {
        pme_gather_special<ORDER>(d, r, s, x, L, kind);
        return;
    }
    int fl[3];
    float th[3][ORDER], dth[3][ORDER];
    const double invL[3] = {1.0 / (double)L.x, 1.0 / (double)L.y, 1.0 / (double)L.z};
    pme_site_setup_f<ORDER, true>(d, x, invL, fl, th, dth);
    const int st = pme_blend_stride(d.gz, ORDER);
    const float *phib = d.pme_blend + (size_t)r * d.gx * d.gy * st;
    const int p0 = fl[2] - (ORDER - 1) + pme_tile_zpad(ORDER);   // first element of the z support in a padded row: >= 0
    const int zb = p0 & ~3, sh = p0 & 3;
    float wz[4 * NV], wdz[4 * NV];
#pragma unroll
    for (int k = 0; k < 4 * NV; k++) {   // wz[k] = th[2][k - sh], zero outside the support (compile-time register indices)
        float t = 0.f, u = 0.f;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (k - q >= 0 && k - q < ORDER) {
                t = sh == q ? th[2][k - q] : t;
                u = sh == q ? dth[2][k - q] : u;
            }
        wz[k] = t;
        wdz[k] = u;
    }
    float fx = 0.f, fy = 0.f, fz = 0.f;
#pragma unroll
    for (int a = 0; a < ORDER; a++) {
        const int ia = pme_wrap(fl[0] - (ORDER - 1) + a, d.gx);
#pragma unroll
        for (int b = 0; b < ORDER; b++) {
            const int ib = pme_wrap(fl[1] - (ORDER - 1) + b, d.gy);
            const float4 *p = reinterpret_cast<const float4 *>(phib + ((size_t)ia * d.gy + ib) * st + zb);
            float s0 = 0.f, sz = 0.f;
#pragma unroll
            for (int v = 0; v < NV; v++) {
                if (NV == 3 && v == 2 && zb + 8 >= st) break;   // the third float4 would lie beyond the row: its weights are zero
                const float4 t = __ldg(p + v);
                s0 = fmaf(wz[4 * v], t.x, s0); s0 = fmaf(wz[4 * v + 1], t.y, s0); s0 = fmaf(wz[4 * v + 2], t.z, s0); s0 = fmaf(wz[4 * v + 3], t.w, s0);
                sz = fmaf(wdz[4 * v], t.x, sz); sz = fmaf(wdz[4 * v + 1], t.y, sz); sz = fmaf(wdz[4 * v + 2], t.z, sz); sz = fmaf(wdz[4 * v + 3], t.w, sz);
            }
            fx = fmaf(dth[0][a] * th[1][b], s0, fx);
            fy = fmaf(th[0][a] * dth[1][b], s0, fy);
            fz = fmaf(th[0][a] * th[1][b], sz, fz);
        }
    }
    const double q = (double)x.w;
    bufC[s] = c0 + (unsigned long long)__double2ll_rn(-q * d.gx / (double)L.x * FORCE_SCALE * (double)fx);
    bufC[cs + s] = c1 + (unsigned long long)__double2ll_rn(-q * d.gy / (double)L.y * FORCE_SCALE * (double)fy);
    bufC[2 * cs + s] = c2 + (unsigned long long)__double2ll_rn(-q * d.gz / (double)L.z * FORCE_SCALE * (double)fz);
}

}  // namespace atm
