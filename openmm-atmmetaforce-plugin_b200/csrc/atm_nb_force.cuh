// atm_nb_force.cuh -- Tier 2 per-step kernels: the two-state direct-space force kernel (nb2_kernel: work items of the
// pair lists, then excluded / 1-4 exception pairs) and the fused scalar stage + merge (nb_merge_kernel).
// Included by atm_nb.cu only.
#pragma once

#include "atm_common.cuh"
#include "atm_nb_types.cuh"

namespace atm {

// ------------------------------------------------------------------------------------------------
// The force kernel.
// ------------------------------------------------------------------------------------------------
// erfc(x) = t (a0 + a1 t + ... + a6 t^6) exp(-x^2), t = 1/(1 + p x); max relative error 6.8e-8 on [0, 4.2]
// (fit: DESIGN.md "erfc"); evaluated in fp32 the rounding error (~4e-7) dominates.
#define ERFC_P 0.357431514f
#define ERFC_A0 1.9957954259e-01f
#define ERFC_A1 2.2717719300e-01f
#define ERFC_A2 5.6467497521e-02f
#define ERFC_A3 5.3674605718e-01f
#define ERFC_A4 -4.8585730230e-01f
#define ERFC_A5 6.4534500206e-01f
#define ERFC_A6 -1.7945805777e-01f

// MUFU wrappers without the denormal / range fix-up code the CUDA math library adds around them
__device__ __forceinline__ float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// round to nearest integer for |x| < 2^22 without the (quarter-rate) FRND instruction
__device__ __forceinline__ float fast_rint(float x) { return __fadd_rn(__fadd_rn(x, 12582912.0f), -12582912.0f); }

struct PairConst {
    float cutoff2;
    float p_alpha;      // ERFC_P * alpha
    float neg_a2_log2e; // -alpha^2 log2(e)
    float two_a_sqrtpi; // 2 alpha / sqrt(pi)
};

// One pair.  qq = q_i q_j k_e (charges are stored pre-multiplied by sqrt(k_e)); sig = (s_i+s_j)/2; eps4 = 4 sqrt(e_i e_j).
// Returns F/r (fscale) and, when ENERGY, the pair energy.
template <bool ENERGY>
__device__ __forceinline__ float pair_interaction(float r2, float qq, float sig, float eps4, const PairConst &pc, float &energy) {
    float rinv = mufu_rsqrt(r2);
#ifdef ATM_RSQRT_NEWTON  // A/B switch: measured force/energy parity is identical without the Newton step
    rinv = rinv * fmaf(-0.5f * r2, rinv * rinv, 1.5f);
#endif
    const float rinv2 = rinv * rinv;
    const float r = r2 * rinv;
    const float s2 = sig * sig * rinv2;
    const float s6 = s2 * s2 * s2;
    const float es6 = eps4 * s6;
    const float flj = es6 * fmaf(12.0f, s6, -6.0f);
    const float t = mufu_rcp(fmaf(pc.p_alpha, r, 1.0f));
    const float ex = mufu_ex2(pc.neg_a2_log2e * r2);
    float poly = fmaf(ERFC_A6, t, ERFC_A5);
    poly = fmaf(poly, t, ERFC_A4);
    poly = fmaf(poly, t, ERFC_A3);
    poly = fmaf(poly, t, ERFC_A2);
    poly = fmaf(poly, t, ERFC_A1);
    poly = fmaf(poly, t, ERFC_A0);
    // Coulomb: E = qq exp(-a^2 r^2) P(t) t / r,  F r = E + qq (2a/sqrt(pi)) exp(-a^2 r^2)  -- five instructions
    const float g = poly * t * rinv;
    const float qe = qq * ex;
    const float ec = qe * g;
    const float fc = fmaf(qe, pc.two_a_sqrtpi, ec);
    if (ENERGY) energy = fmaf(es6, s6, -es6) + ec;
    return (flj + fc) * rinv2;
}

struct ItemCtx {
    int r, A, target, nst;
    const unsigned int *list;
    size_t rsite, comp_stride;
};

template <bool ENERGY, bool STATS>
__device__ __forceinline__ void nb2_item_epilogue(const NbDev &d, const ItemCtx &it, int lane, unsigned long long *buf,
                                                  float (&fix)[CL], float (&fiy)[CL], float (&fiz)[CL], double e_acc, int npairs) {
    // transpose-reduce the 24 i-force accumulators: after three halving exchanges lane (l&7) owns atom l&7
    {
        const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
        float w[4][3];
#pragma unroll
        for (int mm = 0; mm < 4; mm++) {
            const float kx = b0 ? fix[2 * mm + 1] : fix[2 * mm], sx = b0 ? fix[2 * mm] : fix[2 * mm + 1];
            const float ky = b0 ? fiy[2 * mm + 1] : fiy[2 * mm], sy = b0 ? fiy[2 * mm] : fiy[2 * mm + 1];
            const float kz = b0 ? fiz[2 * mm + 1] : fiz[2 * mm], sz = b0 ? fiz[2 * mm] : fiz[2 * mm + 1];
            w[mm][0] = kx + __shfl_xor_sync(0xffffffffu, sx, 1);
            w[mm][1] = ky + __shfl_xor_sync(0xffffffffu, sy, 1);
            w[mm][2] = kz + __shfl_xor_sync(0xffffffffu, sz, 1);
        }
        float x2[2][3];
#pragma unroll
        for (int mm = 0; mm < 2; mm++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float kk = b1 ? w[2 * mm + 1][c] : w[2 * mm][c], ss = b1 ? w[2 * mm][c] : w[2 * mm + 1][c];
                x2[mm][c] = kk + __shfl_xor_sync(0xffffffffu, ss, 2);
            }
        float y[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float kk = b2 ? x2[1][c] : x2[0][c], ss = b2 ? x2[0][c] : x2[1][c];
            y[c] = kk + __shfl_xor_sync(0xffffffffu, ss, 4);
            y[c] += __shfl_xor_sync(0xffffffffu, y[c], 8);
            y[c] += __shfl_xor_sync(0xffffffffu, y[c], 16);
        }
        // lane l (< 8) now holds atom index (b0 + 2 b1 + 4 b2) = l
        if (lane < CL) {
            const int i = it.A * CL + lane;
            red_add_fixed(buf + i, y[0]);
            red_add_fixed(buf + it.comp_stride + i, y[1]);
            red_add_fixed(buf + 2 * it.comp_stride + i, y[2]);
        }
    }
    // energies: warp sum in double, one fixed-point atomic per warp
    if (ENERGY || STATS) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            if (ENERGY) e_acc += __shfl_xor_sync(0xffffffffu, e_acc, off);
            if (STATS) npairs += __shfl_xor_sync(0xffffffffu, npairs, off);
        }
        if (lane == 0) {
            unsigned long long *ea = d.eacc + (size_t)it.r * EACC_SLOTS;
            if (ENERGY) atomicAdd(ea + it.target, (unsigned long long)__double2ll_rn(e_acc * ENERGY_SCALE));
            if (STATS) atomicAdd(ea + 3 + it.target, (unsigned long long)npairs);
        }
    }
}

// Partner data is staged through a per-lane shared-memory ring with cp.async (prefetch distance 3 steps); the cluster
// atoms are broadcast from shared memory instead of living in 48 registers, which buys a fifth resident block per SM.
constexpr int NB_WARPS = NB_THREADS / 32;
#ifndef ATM_PF_DIST
#define ATM_PF_DIST 3
#endif
#ifndef ATM_RING
#define ATM_RING 4
#endif
constexpr int RING = ATM_RING;        // ring slots per lane (power of two, > PF_DIST)
constexpr int PF_DIST = ATM_PF_DIST;  // coordinate steps in flight

// The list entries of a work item (<= ITEM_STEPS x 128 B, contiguous, streaming from DRAM) are brought into shared
// memory by ONE bulk asynchronous copy (the TMA unit: cp.async.bulk, SASS UBLKCP) that signals an mbarrier, instead of
// a rolling register prefetch of one LDG per step.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_arm(unsigned long long *bar, void *smem, const void *gmem, unsigned int bytes) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(b),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void cp_async16(void *smem, unsigned long long gaddr) {  // gaddr: global state space
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gaddr));
}
__device__ __forceinline__ void cp_async8(void *smem, unsigned long long gaddr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gaddr));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct __align__(128) Nb2Smem {
    unsigned int el[NB_WARPS][ITEM_STEPS][32];  // the item's list entries (bulk copy destination)
    float4 xj[NB_WARPS][RING][32];
#ifdef ATM_NB2_PJ_FLOAT4
    float4 pj[NB_WARPS][RING][32];              // .xy = (sigma/2, 2 sqrt(eps)); float4 stride so that ONE byte offset addresses xj and pj
#else
    float2 pj[NB_WARPS][RING][32];              // (sigma/2, 2 sqrt(eps)): the same ELEMENT offset as xj (8-byte stride: half the
                                                // shared-memory wavefronts of the float4-strided variant, one more shift)
#endif
    // cluster atoms in PAIRS (atoms 2p, 2p+1) for the packed f32x2 inner loop: three 128-bit broadcast reads per pair --
    // (x0,x1,y0,y1), (z0,z1,q0,q1), (hs0,hs1,se0,se1)
    float4 ci[NB_WARPS][CL / 2][3];
    unsigned long long bar[NB_WARPS];           // one mbarrier per warp
};
#ifdef ATM_NB2_PJ_FLOAT4
constexpr int PJ_OFF = NB_WARPS * RING * 32;    // float4 elements between a lane's xj slot and its pj slot
#define ATM_PJ_AT(ring, pring, off) reinterpret_cast<float2 *>((ring) + (off) + PJ_OFF)
#else
#define ATM_PJ_AT(ring, pring, off) ((pring) + (off))
#endif

// Pins a value in a register: the compiler can no longer rematerialise it from kernel arguments / thread ids inside the
// list-step loop (at 96 registers it re-derived the ring, accumulator and site-array addresses in every step: ~35 of
// the loop's 336 instructions).
template <typename T>
__device__ __forceinline__ void pin64(T *&p) { asm volatile("" : "+l"(p)); }
__device__ __forceinline__ void pin64(unsigned long long &v) { asm volatile("" : "+l"(v)); }

// ------------------------------------------------------------------------------------------------
// Packed inner loop (sm_100 only): Blackwell executes two fp32 operations per lane in ONE issued instruction
// (fma/mul/add.rn.f32x2 -> SASS FFMA2 / FMUL2 / FADD2, operands in aligned 64-bit register pairs, the second source
// optionally a broadcast scalar, constants as broadcast immediates).  The kernel is bound by instruction issue, not by
// the FMA pipe (round 1: issue slots 80 % busy, FMA pipe 52 %), so the 8 cluster atoms of a list step are processed as
// 4 PAIRS (atoms 2p, 2p+1 against the lane's partner): every arithmetic instruction of pair_interaction, the distance
// and the force accumulation is issued once per two pair slots.  Only the MUFU ops (rsqrt, rcp, ex2) and the
// cutoff / mask selects stay scalar.  The algebra is that of pair_interaction() with the multiplications folded into
// FMAs where that does not lengthen the dependent chain:
//     fs = (es6 (12 s6 - 6) + fc) rinv^2  ->  h = fma(12, s6, -6); fs = fma(es6, h, fc) * rinv2
// f_j is accumulated with the sign of f_i (dx * fs) and negated once per step.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

template <bool ENERGY>
__device__ __forceinline__ float2 pair_interaction_x2(float2 r2, float2 qq, float2 sig, float2 eps4, const PairConst &pc, float2 &energy) {
    const float2 rinv = f2(mufu_rsqrt(r2.x), mufu_rsqrt(r2.y));
    const float2 rinv2 = __fmul2_rn(rinv, rinv);
    const float2 r = __fmul2_rn(r2, rinv);
    const float2 sr = __fmul2_rn(sig, rinv);
    const float2 s2 = __fmul2_rn(sr, sr);
    const float2 s6 = __fmul2_rn(__fmul2_rn(s2, s2), s2);
    const float2 es6 = __fmul2_rn(eps4, s6);
    const float2 h = __ffma2_rn(s6, f2(12.0f), f2(-6.0f));
    const float2 ta = __ffma2_rn(r, f2(pc.p_alpha), f2(1.0f));
    const float2 t = f2(mufu_rcp(ta.x), mufu_rcp(ta.y));
    const float2 ea = __fmul2_rn(r2, f2(pc.neg_a2_log2e));
    const float2 ex = f2(mufu_ex2(ea.x), mufu_ex2(ea.y));
#ifdef ATM_NB2_ESTRIN  // A/B: Estrin's scheme, dependent depth 4 instead of 6 at the price of two more multiplications
    const float2 t2 = __fmul2_rn(t, t), t4 = __fmul2_rn(t2, t2);
    const float2 p01 = __ffma2_rn(t, f2(ERFC_A1), f2(ERFC_A0)), p23 = __ffma2_rn(t, f2(ERFC_A3), f2(ERFC_A2));
    const float2 p45 = __ffma2_rn(t, f2(ERFC_A5), f2(ERFC_A4));
    const float2 p456 = __ffma2_rn(t2, f2(ERFC_A6), p45), p0123 = __ffma2_rn(p23, t2, p01);
    float2 poly = __ffma2_rn(p456, t4, p0123);
#else
    float2 poly = __ffma2_rn(t, f2(ERFC_A6), f2(ERFC_A5));
    poly = __ffma2_rn(poly, t, f2(ERFC_A4));
    poly = __ffma2_rn(poly, t, f2(ERFC_A3));
    poly = __ffma2_rn(poly, t, f2(ERFC_A2));
    poly = __ffma2_rn(poly, t, f2(ERFC_A1));
    poly = __ffma2_rn(poly, t, f2(ERFC_A0));
#endif
    const float2 g = __fmul2_rn(__fmul2_rn(poly, t), rinv);
    const float2 qe = __fmul2_rn(qq, ex);
    const float2 ec = __fmul2_rn(qe, g);
    const float2 fc = __ffma2_rn(qe, f2(pc.two_a_sqrtpi), ec);
    if (ENERGY) energy = __fadd2_rn(__ffma2_rn(es6, s6, f2(-es6.x, -es6.y)), ec);
    return __fmul2_rn(__ffma2_rn(es6, h, fc), rinv2);
}

template <bool ENERGY, bool STATS>
__device__ __forceinline__ void nb2_item(const NbDev &d, const ItemCtx &it, int lane, int w, Nb2Smem &sm, unsigned int parity,
                                         const PairConst &pc) {
    const float4 L = d.box[it.r], iL = d.invbox[it.r];
    const float4 cA = d.cc[(size_t)it.r * d.Cmax + it.A];
    // per-replica bases of the site arrays, pinned: the gather address of a partner is one IMAD.WIDE from them
    unsigned long long xs_g = (unsigned long long)__cvta_generic_to_global(d.xs + it.rsite);
    unsigned long long par_g = (unsigned long long)__cvta_generic_to_global(d.par + it.rsite);
    pin64(xs_g);
    pin64(par_g);

    // the whole entry list of the item: one bulk copy, in flight while the cluster atoms are staged
    if (lane == 0) bulk_load_arm(&sm.bar[w], &sm.el[w][0][0], it.list, (unsigned)it.nst * 128u);
    // cluster atoms -> shared memory (lanes 0..7), shifted next to the cluster centre, interleaved in pairs
    if (lane < CL) {
        float4 x = __ldg(d.xs + it.rsite + (size_t)it.A * CL + lane);
        const float2 pp = __ldg(d.par + it.rsite + (size_t)it.A * CL + lane);
        x.x -= L.x * fast_rint((x.x - cA.x) * iL.x);
        x.y -= L.y * fast_rint((x.y - cA.y) * iL.y);
        x.z -= L.z * fast_rint((x.z - cA.z) * iL.z);
        float *c = reinterpret_cast<float *>(&sm.ci[w][lane >> 1][0]) + (lane & 1);
        c[0] = x.x; c[2] = x.y; c[4] = x.z; c[6] = x.w; c[8] = pp.x; c[10] = pp.y;
    }
    const unsigned int *elp = &sm.el[w][0][lane];
    float4 *ring = &sm.xj[w][0][lane];  // the lane's rings: slot s of xj at ring[32 s], of pj at pring[32 s]
    float2 *pring = reinterpret_cast<float2 *>(&sm.pj[w][0][lane]);
    (void)pring;
    mbar_wait(&sm.bar[w], parity);
#pragma unroll
    for (int q = 0; q < PF_DIST; q++) {
        if (q < it.nst) {
            const unsigned int eq = elp[q * 32] >> 8;
            cp_async16(ring + q * 32, xs_g + (unsigned long long)eq * 16ull);
            cp_async8(ATM_PJ_AT(ring, pring, q * 32), par_g + (unsigned long long)eq * 8ull);
        }
        cp_async_commit();
    }
    __syncwarp();

    float2 fix[CL / 2], fiy[CL / 2], fiz[CL / 2];
#pragma unroll
    for (int p = 0; p < CL / 2; p++) fix[p] = fiy[p] = fiz[p] = f2(0.f);
    // accumulator base of this (target, replica) and the byte stride between the x / y / z blocks, pinned
    unsigned long long *buf = d.buf + (size_t)it.target * 3 * it.comp_stride + it.rsite;
    unsigned long long buf_g = (unsigned long long)__cvta_generic_to_global(buf);
    unsigned long long cs8 = (unsigned long long)it.comp_stride * 8ull;
    pin64(buf_g);
    pin64(cs8);
    double e_acc = 0.0;
    int npairs = 0;
    // periodic shift of the partner in packed form (x, y) + scalar z
    const float2 ncA_xy = f2(-cA.x, -cA.y), iL_xy = f2(iL.x, iL.y), L_xy = f2(L.x, L.y);
    const float4 *ci = &sm.ci[w][0][0];

#ifdef ATM_NB2_STEP_UNROLL   // A/B: let ptxas schedule several list steps (independent partners) together
    constexpr int STEP_UNROLL = ATM_NB2_STEP_UNROLL;
#pragma unroll STEP_UNROLL
#endif
    for (int st = 0; st < it.nst; st++) {
        cp_async_wait<PF_DIST - 1>();  // the group of step st has landed (groups retire in order)
        const int slot = (st & (RING - 1)) * 32;
        const float4 xjc = ring[slot];
        const float2 pjc = *ATM_PJ_AT(ring, pring, slot);
        const unsigned int e = elp[st * 32];
        // keep PF_DIST steps in flight
        {
            const int sp = st + PF_DIST;
            if (sp < it.nst) {
                const unsigned int e_next = elp[sp * 32] >> 8;
                const int ps = (sp & (RING - 1)) * 32;
                cp_async16(ring + ps, xs_g + (unsigned long long)e_next * 16ull);
                cp_async8(ATM_PJ_AT(ring, pring, ps), par_g + (unsigned long long)e_next * 8ull);
            }
            cp_async_commit();
        }
        const unsigned int m = e;  // low 8 bits: exclusion / padding mask
        // minus the partner's coordinates, shifted next to the cluster centre
        float2 tq = __fmul2_rn(__fadd2_rn(f2(xjc.x, xjc.y), ncA_xy), iL_xy);
        tq = __fadd2_rn(__fadd2_rn(tq, f2(12582912.0f)), f2(-12582912.0f));
        const float2 nxy = __ffma2_rn(L_xy, tq, f2(-xjc.x, -xjc.y));
        const float nxj = nxy.x, nyj = nxy.y;
        const float nzj = fmaf(L.z, fast_rint((xjc.z - cA.z) * iL.z), -xjc.z);
        float2 fjx = f2(0.f), fjy = f2(0.f), fjz = f2(0.f), e_step = f2(0.f);
        bool any = false;
#pragma unroll
        for (int p = 0; p < CL / 2; p++) {
            const float4 cxy = ci[3 * p], czq = ci[3 * p + 1], cps = ci[3 * p + 2];
            const float2 dx = __fadd2_rn(f2(cxy.x, cxy.y), f2(nxj));
            const float2 dy = __fadd2_rn(f2(cxy.z, cxy.w), f2(nyj));
            const float2 dz = __fadd2_rn(f2(czq.x, czq.y), f2(nzj));
            const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
            const bool in0 = (r2.x < pc.cutoff2) && !(m & (1u << (2 * p)));
            const bool in1 = (r2.y < pc.cutoff2) && !(m & (2u << (2 * p)));
            // a pair outside the cutoff (or excluded, or a padding slot) is evaluated at r^2 = 1e30: every term underflows
            // to exactly zero (flush-to-zero MUFU paths, no inf/NaN even for r = 0)
            const float2 r2s = f2(in0 ? r2.x : 1e30f, in1 ? r2.y : 1e30f);
            float2 en = f2(0.f);
            const float2 fs = pair_interaction_x2<ENERGY>(r2s, __fmul2_rn(f2(czq.z, czq.w), f2(xjc.w)), __fadd2_rn(f2(cps.x, cps.y), f2(pjc.x)),
                                                          __fmul2_rn(f2(cps.z, cps.w), f2(pjc.y)), pc, en);
            if (ENERGY) e_step = __fadd2_rn(e_step, en);
            if (STATS) npairs += (in0 ? 1 : 0) + (in1 ? 1 : 0);
            any |= in0 | in1;
            fix[p] = __ffma2_rn(dx, fs, fix[p]); fiy[p] = __ffma2_rn(dy, fs, fiy[p]); fiz[p] = __ffma2_rn(dz, fs, fiz[p]);
            fjx = __ffma2_rn(dx, fs, fjx); fjy = __ffma2_rn(dy, fs, fjy); fjz = __ffma2_rn(dz, fs, fjz);
        }
        if (ENERGY) e_acc += (double)(e_step.x + e_step.y);
#ifdef ATM_NB2_ALWAYS_RED  // A/B: no per-lane "any pair in range" predicate, a zero is added instead
        (void)any;
        {
#else
        if (any) {
#endif
            unsigned long long bj = buf_g + (unsigned long long)(e >> 8) * 8ull;
            red_add_fixed_global(bj, -(fjx.x + fjx.y));
            bj += cs8;
            red_add_fixed_global(bj, -(fjy.x + fjy.y));
            bj += cs8;
            red_add_fixed_global(bj, -(fjz.x + fjz.y));
        }
    }
    cp_async_wait<0>();
    float ax[CL], ay[CL], az[CL];
#pragma unroll
    for (int p = 0; p < CL / 2; p++) {
        ax[2 * p] = fix[p].x; ax[2 * p + 1] = fix[p].y;
        ay[2 * p] = fiy[p].x; ay[2 * p + 1] = fiy[p].y;
        az[2 * p] = fiz[p].x; az[2 * p + 1] = fiz[p].y;
    }
    nb2_item_epilogue<ENERGY, STATS>(d, it, lane, buf, ax, ay, az, e_acc, npairs);
}

// ------------------------------------------------------------------------------------------------
// Excluded pairs (Ewald correction -qq erf(ar)/r, minimum image) and 1-4 exceptions (plain Coulomb + LJ, no image).
// One thread per (replica, pair).  Pairs whose atoms move together go to C, others are evaluated in both states.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add_pair_force(const NbDev &d, int r, int target, int si, int sj, float fx, float fy, float fz) {
    unsigned long long *buf = d.buf + (size_t)target * 3 * d.R * d.Smax + (size_t)r * d.Smax;
    const size_t cs = (size_t)d.R * d.Smax;
    red_add_fixed(buf + si, fx); red_add_fixed(buf + cs + si, fy); red_add_fixed(buf + 2 * cs + si, fz);
    red_add_fixed(buf + sj, -fx); red_add_fixed(buf + cs + sj, -fy); red_add_fixed(buf + 2 * cs + sj, -fz);
}

__device__ __forceinline__ void special_pairs_body(const NbDev &d, int t, int r, const int2 *__restrict__ excl, int n_excl,
                                                   const int2 *__restrict__ exc, const float4 *__restrict__ exc_par, int n_exc) {
    double e_tgt[3] = {0.0, 0.0, 0.0};
    if (t < n_excl + n_exc) {
        const bool is_exc = t >= n_excl;
        const int2 pr = is_exc ? exc[t - n_excl] : excl[t];
        const int ga = d.group_of_atom[pr.x], gb = d.group_of_atom[pr.y];
        const float4 L = d.box[r], iL = d.invbox[r];
        const int nstate = (ga == gb) ? 1 : 2;
        for (int s = 0; s < nstate; s++) {
            const int target = (ga == gb) ? TGT_C : (s == 0 ? TGT_S1 : TGT_S2);
            int si = d.site_slot[(size_t)r * d.U + pr.x], sj = d.site_slot[(size_t)r * d.U + pr.y];
            if (s == 1) {
                if (ga != 0) si = d.site_slot[(size_t)r * d.U + d.N + d.ghost_of_atom[pr.x]];
                if (gb != 0) sj = d.site_slot[(size_t)r * d.U + d.N + d.ghost_of_atom[pr.y]];
            }
            const float4 a = d.xs[(size_t)r * d.Smax + si], b = d.xs[(size_t)r * d.Smax + sj];
            float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
            float fs, en;
            if (!is_exc) {
                dx = wrap_delta(dx, L.x, iL.x); dy = wrap_delta(dy, L.y, iL.y); dz = wrap_delta(dz, L.z, iL.z);
                const float r2 = dx * dx + dy * dy + dz * dz;
                const float qq = a.w * b.w;
                if (r2 > 0.f && qq != 0.f && d.alpha > 0.f) {
                    const float rinv = 1.0f / sqrtf(r2), rr = r2 * rinv, ar = d.alpha * rr;
                    const float erf_ar = erff(ar);
                    en = -qq * rinv * erf_ar;
                    fs = -qq * rinv * (erf_ar - ar * expf(-ar * ar) * 1.1283791670955126f) * rinv * rinv;
                } else { en = 0.f; fs = 0.f; }
            } else {
                const float4 pp = exc_par[t - n_excl];  // ke*chargeProd, sigma, 4 eps
                const float r2 = dx * dx + dy * dy + dz * dz;
                const float rinv = 1.0f / sqrtf(r2), rinv2 = rinv * rinv;
                const float s2 = pp.y * pp.y * rinv2, s6 = s2 * s2 * s2;
                en = pp.z * s6 * (s6 - 1.0f) + pp.x * rinv;
                fs = (pp.z * s6 * (12.0f * s6 - 6.0f) + pp.x * rinv) * rinv2;
            }
            add_pair_force(d, r, target, si, sj, dx * fs, dy * fs, dz * fs);
            e_tgt[target] += (double)en;
        }
    }
    // block reduction of the three energies (warp shuffle, then one atomic per warp)
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double v = e_tgt[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0 && v != 0.0) {
            unsigned long long *ea = d.eacc + (size_t)r * EACC_SLOTS;
            atomicAdd(ea + k, (unsigned long long)__double2ll_rn(v * ENERGY_SCALE));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fused scalar stage + merge.  Per replica: u = U(S2) - U(S1), soft-core, softplus, sp (double, once per block),
// then  F[slot] += C + (1-sp) S1 + sp S2  gathered from cluster order, accumulators zeroed behind the read.
// Semantics: CommonATMMetaForceKernels.cpp:182-201 + kernels/atmmetaforce.cc:8-16, blend in double.
// ------------------------------------------------------------------------------------------------
constexpr int MERGE2_THREADS = 256;

// Scalar stage, one thread per replica: u = U(S2) - U(S1), soft-core, softplus, sp -- all in double on the device
// (the reference does this on the host after two blocking energy downloads, CommonATMMetaForceKernels.cpp:164-199).
__device__ double scalar_stage_replica(const NbDev &d, int r, const double *__restrict__ energy_ext, int include_energy,
                                       bool write_record) {
    volatile unsigned long long *ea = d.eacc + (size_t)r * EACC_SLOTS;  // written by atomics of other blocks: read through L2
    const double uc = (double)(long long)ea[0] / ENERGY_SCALE, u1 = (double)(long long)ea[1] / ENERGY_SCALE,
                 u2 = (double)(long long)ea[2] / ENERGY_SCALE;
    double U1 = uc + u1, U2 = uc + u2, du = u2 - u1;
    double *e = d.energies + (size_t)r * ATM_NUM_ENERGY_SLOTS;
    double rec1 = 0.0, rec2 = 0.0, eself = 0.0;
    if (d.pme_on) {
        // reciprocal energies of the two states (accumulated in double, so their difference is as good as the sum)
        const double r1 = (double)(long long)ea[6] / ENERGY_SCALE, r2 = (double)(long long)ea[7] / ENERGY_SCALE;
        const double self = -d.pme_self_sum * (double)d.alpha * 0.5641895835477563;  // -alpha/sqrt(pi) sum q^2
        const float4 Lb = d.box[r];
        const double bg = -3.141592653589793 * d.pme_qtot2 /
                          (2.0 * (double)Lb.x * (double)Lb.y * (double)Lb.z * (double)d.alpha * (double)d.alpha);
        U1 += r1 + self + bg;
        U2 += r2 + self + bg;
        du += r2 - r1;
        rec1 = r1; rec2 = r2; eself = self;
    }
    if (d.disp_coeff != 0.0) {  // same constant in both states: u is unaffected
        const float4 Lb = d.box[r];
        const double ed = d.disp_coeff / ((double)Lb.x * (double)Lb.y * (double)Lb.z);
        U1 += ed;
        U2 += ed;
    }
    if (energy_ext) {
        U1 += energy_ext[2 * r];
        U2 += energy_ext[2 * r + 1];
        du += energy_ext[2 * r + 1] - energy_ext[2 * r];
    }
    Scalars s = scalar_stage(d.params + (size_t)r * ATM_NUM_PARAMS, U1, U2, du);
    if (d.flags[0] & 15) {
        // the pair lists of the last rebuild are incomplete (a list outgrew its capacity, or the box is too small for
        // the list radius), or a site has moved further since that rebuild than the structure tolerates (bit 3, set by
        // the PME gather): nothing computed from them may look like a result
        const double bad = __longlong_as_double(0x7ff8000000000000ll);
        U1 = U2 = bad;
        s.u = s.usc = s.ebias = s.energy = s.sp = bad;
    }
    if (write_record) {
        e[ATM_E_UREC1] = rec1; e[ATM_E_UREC2] = rec2; e[ATM_E_USELF] = eself;
        e[ATM_E_U1] = U1; e[ATM_E_U2] = U2; e[ATM_E_U] = s.u; e[ATM_E_USC] = s.usc; e[ATM_E_EBIAS] = s.ebias;
        e[ATM_E_ENERGY] = include_energy ? s.energy : 0.0; e[ATM_E_SP] = s.sp;
        e[ATM_E_NPAIRS] = (double)(ea[3] + ea[4] + ea[5]);
        e[ATM_E_NPAIRS_C] = (double)ea[3]; e[ATM_E_NPAIRS_S1] = (double)ea[4]; e[ATM_E_NPAIRS_S2] = (double)ea[5];
    }
    return s.sp;  // the accumulators are zeroed by the next step's pack kernel
}

// ------------------------------------------------------------------------------------------------
// The per-step compute launch: work items of the pair lists (one warp each), then blocks of excluded / exception
// pairs.
// ------------------------------------------------------------------------------------------------
struct SpecialArgs {
    const int2 *excl;
    const int2 *exc;
    const float4 *exc_par;
    int n_excl, n_exc;
    int blocks_per_replica;  // ceil((n_excl + n_exc) / NB_THREADS)
    int special_first;       // 1: the special-pair blocks are the first blocks of the grid, 0: the last (round 1)
};

template <bool STATS>
__global__ void __launch_bounds__(NB_THREADS, NB_MIN_BLOCKS)
nb2_kernel(NbDev d, int n_item_blocks, int energy_common, SpecialArgs sp) {
    pdl_trigger();  // the merge may be scheduled once every block of this grid has started (it waits for completion)
    pdl_wait();     // cluster-order coordinates come from the pack kernel
    const int lane = threadIdx.x & 31;
    // the special-pair blocks are the LAST blocks of the grid by default (sp.special_first: A/B switch, measured slower)
    const int n_special = sp.blocks_per_replica * d.R;
    const int bid = sp.special_first ? (int)blockIdx.x - n_special : (int)blockIdx.x;
    if (bid >= 0 && bid < n_item_blocks) {
        // The pruned list's item count lives on the device; the grid is sized from the count the last verified build
        // saw plus a margin, and this grid-stride loop picks up whatever a later prune added beyond it.
        __shared__ Nb2Smem sm;
        const int w = threadIdx.x >> 5;
        if (lane == 0) mbar_init(&sm.bar[w], 1);
        __syncwarp();
        unsigned int parity = 0;  // phase of this warp's mbarrier: flips with every completed bulk copy
        PairConst pc;
        pc.cutoff2 = d.cutoff2;
        pc.p_alpha = ERFC_P * d.alpha;
        pc.neg_a2_log2e = -d.alpha * d.alpha * 1.4426950408889634f;
        pc.two_a_sqrtpi = d.two_alpha_over_sqrtpi;
        // Work items live in ITEM_STEPS length buckets and are handed out longest first.  Lane l keeps the inclusive
        // prefix count of the buckets ITEM_STEPS, ITEM_STEPS-1, ..., ITEM_STEPS-l (one load + a warp scan per WARP);
        // an item index is then turned into (bucket, index in bucket) by a ballot instead of a chain of up to 16
        // dependent loads per ITEM.
        const int cnt = (lane < ITEM_STEPS) ? d.iflags[ITEM_BUCKET0 + ITEM_STEPS - lane] : 0;
        int pre = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, pre, off);
            if (lane >= off) pre += v;
        }
        const int n_items = d.iflags[4], stride = n_item_blocks * (NB_THREADS / 32);
        for (int warp = bid * (NB_THREADS / 32) + (threadIdx.x >> 5); warp < n_items; warp += stride, parity ^= 1u) {
            __syncwarp();  // the previous item's readers of this warp's shared-memory slots are done
            const unsigned int below = __ballot_sync(0xffffffffu, warp < pre);  // lanes whose prefix covers this item
            const int l = below ? __ffs(below) - 1 : ITEM_STEPS - 1;
            const int b = ITEM_STEPS - l;
            const int wi = warp - __shfl_sync(0xffffffffu, pre - cnt, l);
            const int4 item = __ldg(d.items + (size_t)b * d.max_items + wi);
            ItemCtx it;
            it.r = item.z & 0xff;
            it.A = item.y & 0x0fffffff;
            it.target = (item.y >> 28) & 3;
            it.nst = item.z >> 8;
            it.list = d.jlist + (unsigned int)item.x;
            it.rsite = (size_t)it.r * d.Smax;
            it.comp_stride = (size_t)d.R * d.Smax;
            if (it.target == TGT_C && !energy_common) nb2_item<false, STATS>(d, it, lane, w, sm, parity, pc);
            else nb2_item<true, STATS>(d, it, lane, w, sm, parity, pc);
        }
    } else {
        const int sb = sp.special_first ? (int)blockIdx.x : (int)blockIdx.x - n_item_blocks;
        const int r = sb / sp.blocks_per_replica, chunk = sb - r * sp.blocks_per_replica;
        special_pairs_body(d, chunk * NB_THREADS + threadIdx.x, r, sp.excl, sp.n_excl, sp.exc, sp.exc_par, sp.n_exc);
    }
}

// Merge, one thread per cluster-order slot: the three accumulators are read (and zeroed) coalesced, only the final
// read-modify-write of the caller's force buffer is a scatter.  The thread of a displaced atom also folds in (and
// zeroes) the S2 accumulator of its ghost site; ghost and padding slots have no thread work.
// F[slot] += C + llrint(sp * S2 + (1 - sp) * S1), blend in double (kernels/atmmetaforce.cc:8-16 semantics).
// The scalar stage is fused in: thread 0 of every block forms u, the soft core, the softplus bias and sp = dW/du of the
// block's replica in double (the reference does this on the host after two blocking energy downloads,
// CommonATMMetaForceKernels.cpp:164-199); every block gets the same bits, block 0 of a replica writes the energy record.
__global__ void __launch_bounds__(MERGE2_THREADS)
nb_merge_kernel(NbDev d, long long *__restrict__ force, const long long *__restrict__ f1_ext,
                const long long *__restrict__ f2_ext, const double *__restrict__ energy_ext, int include_energy) {
    __shared__ double s_sp;
    const int r = blockIdx.y;
    const int s = blockIdx.x * MERGE2_THREADS + threadIdx.x;
    const size_t rsite = (size_t)r * d.Smax;
    // everything that does not depend on the force kernel first
    const bool live = s < CL * d.nclusters[r];
    const int i = live ? d.slot_out[rsite + s] : -1;  // caller's slot of this site's atom, -1 for ghosts and padding
    const int gs = i >= 0 ? d.slot_ghost[rsite + s] : -1;
    pdl_wait();
    if (threadIdx.x == 0) s_sp = scalar_stage_replica(d, r, energy_ext, include_energy, blockIdx.x == 0);
    __syncthreads();
    if (i < 0) return;
    const double sp = s_sp, sp1 = 1.0 - sp;
    const size_t cs = (size_t)d.R * d.Smax;
    long long *bufC = (long long *)d.buf + rsite, *buf1 = bufC + 3 * cs, *buf2 = bufC + 6 * cs;
    long long fc[3], fa[3], fb[3], fg[3], fo_old[3];
    size_t fo[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {  // every load first
        fc[c] = bufC[c * cs + s];
        fa[c] = buf1[c * cs + s];
        fb[c] = buf2[c * cs + s];
        fg[c] = gs >= 0 ? buf2[c * cs + gs] : 0;
        fo[c] = (size_t)r * 3 * d.P + (size_t)c * d.P + i;
        fo_old[c] = force[fo[c]];
    }
    bool nz1[3], nz2[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        nz1[c] = fa[c] != 0;
        nz2[c] = fb[c] != 0;
        if (f1_ext) fa[c] += f1_ext[fo[c]];
        if (f2_ext) fb[c] += f2_ext[fo[c]];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        // hand the accumulators back zeroed; the state-specific ones are zero already for every site that is not
        // within the cutoff of a displaced atom or a ghost (96 % of them): skip those writes
        bufC[c * cs + s] = 0;
        if (nz1[c]) buf1[c * cs + s] = 0;
        if (nz2[c]) buf2[c * cs + s] = 0;
        if (gs >= 0) buf2[c * cs + gs] = 0;
        const double v = __dadd_rn(__dmul_rn(sp, (double)(fb[c] + fg[c])), __dmul_rn(sp1, (double)fa[c]));
        force[fo[c]] = fo_old[c] + fc[c] + __double2ll_rn(v);
    }
}

}  // namespace atm
