/*
 * atm_oracle.h -- CPU restatement (plain C, double precision) of the ATM Meta-Force per-step hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (openmm-atmmetaforce-plugin_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: the ATM-specific algebra (copy-state, soft-core, softplus, merge) restates the
 * reference's Reference-platform and common-platform kernels line by line (citations at every
 * function).  The NonbondedForce arithmetic behind U1/U2 lives in OpenMM (>= 7.7, no pinned version;
 * reference README.md:32), which is NOT in /root/reference and not installable here; it is restated
 * from OpenMM's published theory (LJ 12-6 Lorentz-Berthelot + Ewald real space inside the cutoff,
 * erf exclusion correction, 1-4 exceptions, Ewald reciprocal).  The only golden vector the reference
 * holds for this path, u = 58.2 +- 0.1 kJ/mol on the TEMOA-G1 fixture
 * (reference python/tests/test_abfe.py:148,150), is reproduced (58.23) -- see tests/test_oracle_pin.py.
 * Forces, the RBFE system, the soft-core branch, alpha > 0 and direction = -1 are PARITY UNPINNED by
 * the reference (it has no test for them); they are covered by self-consistency tests only.
 */
#ifndef ATM_ORACLE_H_
#define ATM_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATM_ORACLE_ONE_4PI_EPS0 138.935456 /* kJ nm / (mol e^2), OpenMM 7.x constant */

/* ---- scalar stage: reference platforms/reference/src/ReferenceATMMetaForceKernels.cpp:26-37, 82-98;
 *      identical math in platforms/common/src/CommonATMMetaForceKernels.cpp:19-30, 182-199 ---- */
double atm_oracle_softcore(double u, double umax, double a, double ub, double *fp);

/* p = {lambda1, lambda2, alpha, u0, w0, umax, ubcore, acore, direction}
 * out = {u_sc (perturbation energy), fp, ebias, bfp, energy (= e0 + ebias), sp_common, sp_reference}
 *   sp_common    = dir > 0 ? bfp*fp : 1 - bfp*fp   (weight of F2, CommonATMMetaForceKernels.cpp:199, kept in double)
 *   sp_reference = bfp*fp                           (ReferenceATMMetaForceKernels.cpp:101) */
void atm_oracle_scalars(const double p[9], double U1, double U2, double out[7]);

/* ---- displacement table: CommonATMMetaForceKernels.cpp:83-106 (entry i of the force == atom i;
 *      slot s of the device order holds atom atom_index[s]; d rounded double -> float; padded tail zero) ---- */
void atm_oracle_displ_table(int n, int padded, const int32_t *atom_index, const double *dxyz, float *displ4);

/* ---- copy-state ----
 * f32: common kernel in single / mixed precision (kernels/atmmetaforce.cc:33-51): real = float.
 * f64: common kernel in double precision: real = double, displacement still the float table.
 * ref: Reference platform, everything double (ReferenceATMMetaForceKernels.cpp:116-124). */
void atm_oracle_copy_state_f32(int n, const float *posq, const float *corr, const float *displ4,
                               float *posq1, float *corr1, float *posq2, float *corr2);
void atm_oracle_copy_state_f64(int n, const double *posq, const float *displ4, double *posq1, double *posq2);
void atm_oracle_copy_state_ref(int n, const double *pos, const double *displ, double *pos1, double *pos2);

/* ---- merge ----
 * ref: double forces, Reference platform :99-110 (direction handled by swapping the roles).
 * i64: the 2^32 fixed-point SoA buffers of the GPU platforms (kernels/atmmetaforce.cc:1-17) with the blend
 *      carried out in double and rounded to nearest: force[c*P+i] += llrint(sp*f2 + (1-sp)*f1). */
void atm_oracle_merge_ref(int n, double *force, const double *f1, const double *f2, double sp_reference,
                          double direction);
void atm_oracle_hybrid_force_i64(int n, int padded, int64_t *force, const int64_t *f1, const int64_t *f2, double sp);

/* ---- NonbondedForce direct space (what the inner contexts evaluate, ATMMetaForceImpl.cpp:113,116) ---- */
typedef struct {
    int n;
    const double *charge;   /* e */
    const double *sigma;    /* nm */
    const double *epsilon;  /* kJ/mol */
    int n_excl;
    const int32_t *excl;    /* [n_excl][2] excluded pairs (each once) */
    int n_exc14;
    const int32_t *exc14;   /* [n_exc14][2] exception pairs (subset of excl) */
    const double *exc14_par;/* [n_exc14][3] chargeProd, sigma, epsilon */
    double box[3];          /* rectangular */
    double cutoff;          /* nm */
    double ewald_alpha;     /* 1/nm */
} atm_oracle_system;

/* energies out[4] = {lj, coulomb real-space, exclusion correction, exceptions}; returns their sum.
 * force (3n, kJ/mol/nm) is accumulated (+=) when non-NULL.  OpenMP parallel. */
double atm_oracle_nb_direct(const atm_oracle_system *sys, const double *pos, double *force, double out[4]);

/* Exact Ewald reciprocal sum (k-space cut at exp(-k^2/4a^2) < tol); self energy NOT included.
 * force accumulated when non-NULL. */
double atm_oracle_ewald_recip(const atm_oracle_system *sys, const double *pos, double tol, double *force);

/* Smooth particle-mesh Ewald reciprocal energy (Essmann et al. 1995; the algorithm OpenMM's NonbondedForce uses for
 * PME): B-splines of the given order (4..8), grid (n[0], n[1], n[2]), plain separable DFTs in double.  Self energy NOT
 * included.  force accumulated when non-NULL.  With a fine grid it converges to atm_oracle_ewald_recip. */
double atm_oracle_pme_recip(const atm_oracle_system *sys, const double *pos, const int n[3], int order, double *force);

/* Whole reference step on the CPU: copy-state (ref), two full direct-space evaluations, scalar stage, merge.
 * pos, displ: 3n doubles; force_out (3n) is accumulated into; energies[5] = {U1, U2, u_sc, energy, sp_common}.
 * du_ext is added to U2 (e.g. the reciprocal-space difference computed elsewhere). */
void atm_oracle_step(const atm_oracle_system *sys, const double p[9], const double *pos, const double *displ,
                     double du_ext, double *force_out, double energies[5]);

int atm_oracle_num_threads(void);
void atm_oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
