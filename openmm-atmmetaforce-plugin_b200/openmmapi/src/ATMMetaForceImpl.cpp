// ATMMetaForceImpl.cpp -- see ATMMetaForceImpl.h.  Host logic only; every number comes from libatm_b200.so.
#include "ATMMetaForceImpl.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "ATMMetaForceB200Kernel.h"
#include "ATMMetaForceKernels.h"

using namespace ATMMetaForcePlugin;
using OpenMM::OpenMMException;

#ifndef ATM_HAVE_OPENMM

namespace {

void check(int rc, const char *what) {
    if (rc != ATM_OK) throw OpenMMException(std::string(what) + ": " + atm_last_error());
}

// largest displacement of any particle since `ref` (upper bound: sqrt(3) * largest component)
double maxMove(const std::vector<OpenMM::Vec3> &pos, const std::vector<OpenMM::Vec3> &ref) {
    double m = 0.0;
    for (size_t i = 0; i < pos.size(); i++)
        for (int c = 0; c < 3; c++) m = std::max(m, std::fabs(pos[i][c] - ref[i][c]));
    return m * 1.7320508075688772;
}

}  // namespace

ATMMetaForceImpl::ATMMetaForceImpl(const ATMMetaForce &owner)
    : innerIntegrator1(1.0), innerIntegrator2(1.0), hasInitializedInnerContexts(false), owner(owner), nonbonded(nullptr),
      PerturbationEnergy(0.0), variable_force_groups_mask(0), device(-1), skin(0.05),
      skinOuter(0.3), handle(nullptr), pipeline(nullptr), stream(nullptr), posqHost(nullptr), forceHost(nullptr),
      energyHost(nullptr), extHost{nullptr, nullptr}, energyExtHost(nullptr), paddedNumAtoms(0), displacementsDirty(false), reciprocalOn(false), boxVersionSeen(0),
      energyRecord(ATM_NUM_ENERGY_SLOTS, 0.0) {}

ATMMetaForceImpl::~ATMMetaForceImpl() {
    innerContext1.reset();
    innerContext2.reset();
    releaseBackend();
}

void ATMMetaForceImpl::copysystem(const OpenMM::System &system, OpenMM::System &innerSystem) {
    for (int i = 0; i < system.getNumParticles(); i++) innerSystem.addParticle(system.getParticleMass(i));
    for (int i = 0; i < system.getNumConstraints(); i++) {
        int p1, p2;
        double distance;
        system.getConstraintParameters(i, p1, p2, distance);
        innerSystem.addConstraint(p1, p2, distance);
    }
    OpenMM::Vec3 a, b, c;
    system.getDefaultPeriodicBoxVectors(a, b, c);
    innerSystem.setDefaultPeriodicBoxVectors(a, b, c);
    // every Force outside the ATM force group goes into the inner systems, keeping its group; which of them are
    // evaluated is decided per call by the variable-force-group mask.  A NonbondedForce keeps its reciprocal space in
    // its own group.
    for (int i = 0; i < system.getNumForces(); i++) {
        const OpenMM::Force &force = system.getForce(i);
        if (force.getForceGroup() == owner.getForceGroup()) continue;
        OpenMM::Force *copy = force.clone();
        if (!copy) throw OpenMMException("ATMMetaForce: a Force of the System cannot be cloned into the inner contexts");
        copy->setForceGroup(force.getForceGroup());
        if (auto *nb = dynamic_cast<OpenMM::NonbondedForce *>(copy)) nb->setReciprocalSpaceForceGroup(-1);
        innerSystem.addForce(copy);
    }
}

void ATMMetaForceImpl::releaseBackend() {
    if (pipeline) atm_host_pipeline_destroy(pipeline);
    if (handle) atm_destroy(handle);
    if (posqHost) atm_host_free(posqHost);
    if (forceHost) atm_host_free(forceHost);
    if (energyHost) atm_host_free(energyHost);
    if (extHost[0]) atm_host_free(extHost[0]);
    if (extHost[1]) atm_host_free(extHost[1]);
    if (energyExtHost) atm_host_free(energyExtHost);
    if (stream) atm_stream_destroy(stream);
    pipeline = nullptr; handle = nullptr; posqHost = nullptr; forceHost = nullptr; energyHost = nullptr; stream = nullptr;
    extHost[0] = extHost[1] = nullptr; energyExtHost = nullptr;
}

void ATMMetaForceImpl::setPairListSkins(double inner_nm, double outer_nm) {
    if (!(inner_nm > 0.0) || outer_nm < inner_nm) throw OpenMMException("ATMMetaForce: pair-list skins must satisfy 0 < inner <= outer");
    if (handle) throw OpenMMException("ATMMetaForce: the pair-list skins must be set before the first evaluation");
    skin = inner_nm;
    skinOuter = outer_nm;
}

void ATMMetaForceImpl::initialize(OpenMM::ContextImpl &context) {
    const OpenMM::System &system = context.getSystem();
    variable_force_groups_mask = variableForceGroupsMask(owner);   // throws when the ATM group itself is listed
    if (context.getPlatform().supportsKernels(getKernelNames())) {
        // a platform back-end is registered: the reference's own orchestration (ref: initialize :83-87)
        copysystem(system, innerSystem1);
        copysystem(system, innerSystem2);
        kernel = context.getPlatform().createKernel(CalcATMMetaForceKernel::Name(), context);
        kernel.getAs<CalcATMMetaForceKernel>().initialize(system, owner);
        return;
    }
    if (owner.getNumParticles() != system.getNumParticles())
        throw OpenMMException("ATMMetaForce must have exactly as many particles as the System it belongs to.");
    // which Forces the two states evaluate: everything outside the ATM group that sits in a variable force group
    // (ref: copysystem :51-65 clones every non-ATM force; the mask :75-81 selects the variable ones at evaluation time)
    nonbonded = nullptr;
    hostForces.clear();
    for (int i = 0; i < system.getNumForces(); i++) {
        const OpenMM::Force &f = system.getForce(i);
        if (&f == &owner || f.getForceGroup() == owner.getForceGroup()) continue;
        if (!((variable_force_groups_mask >> f.getForceGroup()) & 1)) continue;
        if (auto *hf = dynamic_cast<const OpenMM::HostEvaluatedForce *>(&f)) {
            hostForces.push_back(hf);   // evaluated at both states, fed to the step as the external per-state terms
            continue;
        }
        const OpenMM::NonbondedForce *nb = dynamic_cast<const OpenMM::NonbondedForce *>(&f);
        if (!nb) throw OpenMMException("ATMMetaForce: a variable force group holds a Force this OpenMM-free build cannot evaluate "
                                       "(NonbondedForce and host-evaluated forces only)");
        if (nonbonded) throw OpenMMException("ATMMetaForce: more than one NonbondedForce in the variable force groups");
        if (nb->getNumParticles() != system.getNumParticles())
            throw OpenMMException("NonbondedForce must have exactly as many particles as the System it belongs to.");
        if (nb->getNonbondedMethod() != OpenMM::NonbondedForce::PME && nb->getNonbondedMethod() != OpenMM::NonbondedForce::Ewald)
            throw OpenMMException("ATMMetaForce: this back-end evaluates periodic Ewald / PME direct space only");
        nonbonded = nb;
    }
}

std::map<std::string, double> ATMMetaForceImpl::getDefaultParameters() {
    return ATMMetaForceB200Kernel::getDefaultParameters(owner);
}

std::vector<std::string> ATMMetaForceImpl::getKernelNames() { return {ATMMetaForceB200Kernel::Name()}; }

void ATMMetaForceImpl::updateParametersInContext(OpenMM::ContextImpl &context) {
    if (owner.getNumParticles() != context.getSystem().getNumParticles())
        throw OpenMMException("copyParametersToContext: The number of ATMMetaForce particles has changed");
    if (kernel) {
        kernel.getAs<CalcATMMetaForceKernel>().copyParametersToContext(context, owner);   // ref: :150-152
        return;
    }
    displacementsDirty = true;   // uploaded (and the pair lists rebuilt) by the next evaluation
}

void ATMMetaForceImpl::createBackend(OpenMM::ContextImpl &context) {
    if (!nonbonded)
        throw OpenMMException("ATMMetaForce: no NonbondedForce in the variable force groups: nothing to evaluate");
    const int n = owner.getNumParticles();
    try {
        atm_config cfg;
        cfg.num_particles = n;
        cfg.padded_num_particles = 0;
        cfg.precision = ATM_PREC_MIXED;
        cfg.num_replicas = 1;
        cfg.device = device;
        check(atm_create(&cfg, &handle), "ATMMetaForce: creating the Blackwell back-end");
        paddedNumAtoms = 32 * ((n + 31) / 32);
        check(atm_stream_create(device, &stream), "ATMMetaForce: stream");
        displacements = owner.getDisplacementArray();
        check(atm_set_displacements(handle, nullptr, displacements.data(), stream), "ATMMetaForce: uploading the displacement table");
        {   // the box goes in before the NonbondedForce description, so that every allocation and upload of the set-up is
            // issued on `stream`
            OpenMM::Vec3 a, b, c;
            context.getPeriodicBoxVectors(a, b, c);
            const double box[9] = {a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]};
            check(atm_set_box(handle, -1, box), "ATMMetaForce: periodic box");
        }
        // NonbondedForce -> atm_nonbonded_desc: every exception excludes its pair; one with a non-zero chargeProd or
        // epsilon is a scaled 1-4 interaction with its own parameters
        std::vector<double> q(n), sig(n), eps(n);
        for (int i = 0; i < n; i++) nonbonded->getParticleParameters(i, q[i], sig[i], eps[i]);
        std::vector<int32_t> excl, excPairs;
        std::vector<double> excParams;
        for (int e = 0; e < nonbonded->getNumExceptions(); e++) {
            int a, b;
            double cp, s, ep;
            nonbonded->getExceptionParameters(e, a, b, cp, s, ep);
            if (a < 0 || a >= n || b < 0 || b >= n || a == b) throw OpenMMException("NonbondedForce: Illegal particle index for an exception");
            excl.push_back(a); excl.push_back(b);
            if (cp != 0.0 || ep != 0.0) {
                excPairs.push_back(a); excPairs.push_back(b);
                excParams.push_back(cp); excParams.push_back(s); excParams.push_back(ep);
            }
        }
        atm_nonbonded_desc desc;
        std::memset(&desc, 0, sizeof(desc));
        desc.charge = q.data(); desc.sigma = sig.data(); desc.epsilon = eps.data();
        desc.num_exclusions = (int32_t)(excl.size() / 2); desc.exclusions = excl.data();
        desc.num_exceptions = (int32_t)(excPairs.size() / 2); desc.exception_pairs = excPairs.data(); desc.exception_params = excParams.data();
        desc.cutoff = nonbonded->getCutoffDistance();
        // OpenMM's rule for the Ewald splitting parameter: alpha = sqrt(-ln(2 tol)) / r_c
        desc.ewald_alpha = std::sqrt(-std::log(2.0 * nonbonded->getEwaldErrorTolerance())) / desc.cutoff;
        desc.skin = skin;
        desc.skin_outer = skinOuter;
        check(atm_nb_setup(handle, &desc, stream), "ATMMetaForce: describing the NonbondedForce to the back-end");
        // The reference's inner contexts evaluate the cloned NonbondedForce as a whole (copysystem puts its reciprocal
        // space into the force's own group, ref: :60-62): reciprocal space (smooth PME on OpenMM's mesh rule, order 5)
        // and the long-range dispersion correction are part of U1 / U2 -- unless the caller moved the reciprocal space
        // into a force group that is not variable, i.e. evaluates it elsewhere.
        const int rg = nonbonded->getReciprocalSpaceForceGroup();
        if (rg < 0 || ((variable_force_groups_mask >> rg) & 1)) {
            OpenMM::Vec3 box[3];
            context.getPeriodicBoxVectors(box[0], box[1], box[2]);
            int grid[3];
            OpenMM::pmeGridDimensions(desc.ewald_alpha, nonbonded->getEwaldErrorTolerance(), box, grid);
            check(atm_pme_setup(handle, grid[0], grid[1], grid[2], 5), "ATMMetaForce: PME mesh");
            reciprocalOn = true;
        }
        check(atm_nb_set_dispersion_correction(handle, nonbonded->getUseDispersionCorrection() ? 1 : 0), "ATMMetaForce: dispersion correction");
        check(atm_host_alloc(sizeof(float) * 4 * (size_t)paddedNumAtoms, (void **)&posqHost), "ATMMetaForce: pinned coordinates");
        check(atm_host_alloc(sizeof(int64_t) * 3 * (size_t)paddedNumAtoms, (void **)&forceHost), "ATMMetaForce: pinned forces");
        check(atm_host_alloc(sizeof(double) * ATM_NUM_ENERGY_SLOTS, (void **)&energyHost), "ATMMetaForce: pinned energy record");
        if (!hostForces.empty()) {
            for (int s = 0; s < 2; s++) {
                check(atm_host_alloc(sizeof(int64_t) * 3 * (size_t)paddedNumAtoms, (void **)&extHost[s]), "ATMMetaForce: pinned external forces");
                std::memset(extHost[s], 0, sizeof(int64_t) * 3 * (size_t)paddedNumAtoms);
            }
            check(atm_host_alloc(sizeof(double) * 2, (void **)&energyExtHost), "ATMMetaForce: pinned external energies");
        }
        std::memset(posqHost, 0, sizeof(float) * 4 * (size_t)paddedNumAtoms);
        for (int i = 0; i < n; i++) posqHost[4 * i + 3] = (float)q[i];
        check(atm_host_pipeline_create(1, &handle, &pipeline), "ATMMetaForce: host pipeline");
        check(atm_stream_synchronize(nullptr), "ATMMetaForce: set-up");   // nothing of the set-up is left in flight on the
        check(atm_stream_synchronize(stream), "ATMMetaForce: set-up");    // legacy stream or on ours
    } catch (...) {
        releaseBackend();
        throw;
    }
    displacementsDirty = false;
    boxVersionSeen = context.getBoxVersion();
    refRebuild.clear();
    refPrune.clear();
}

// The reference's calcForcesAndEnergy (ref: :90-128): inner contexts on first use, copyState, the two inner evaluations
// of the variable force groups, the kernel's scalar stage + hybrid merge.
double ATMMetaForceImpl::calcWithPlatformKernel(OpenMM::ContextImpl &context, bool includeForces, bool includeEnergy, int groups) {
    if (!hasInitializedInnerContexts) {
        hasInitializedInnerContexts = true;
        innerContext1.reset(context.createLinkedContext(innerSystem1, innerIntegrator1));
        innerContext2.reset(context.createLinkedContext(innerSystem2, innerIntegrator2));
        std::vector<OpenMM::Vec3> pos;
        context.getPositions(pos);
        innerContext1->setPositions(pos);
        innerContext2->setPositions(pos);
    }
    if ((groups & (1 << owner.getForceGroup())) == 0) return 0.0;
    OpenMM::ContextImpl &inner1 = OpenMM::getContextImpl(*innerContext1), &inner2 = OpenMM::getContextImpl(*innerContext2);
    CalcATMMetaForceKernel &k = kernel.getAs<CalcATMMetaForceKernel>();
    k.copyState(context, inner1, inner2);
    const double State1Energy = inner1.calcForcesAndEnergy(true, true, variable_force_groups_mask);
    const double State2Energy = inner2.calcForcesAndEnergy(true, true, variable_force_groups_mask);
    const double energy = k.execute(context, inner1, inner2, State1Energy, State2Energy, includeForces, includeEnergy);
    PerturbationEnergy = k.getPerturbationEnergy();
    std::fill(energyRecord.begin(), energyRecord.end(), 0.0);
    energyRecord[ATM_E_U1] = State1Energy;
    energyRecord[ATM_E_U2] = State2Energy;
    energyRecord[ATM_E_U] = context.getParameter(ATMMetaForce::Direction()) > 0 ? State2Energy - State1Energy : State1Energy - State2Energy;
    energyRecord[ATM_E_USC] = PerturbationEnergy;
    energyRecord[ATM_E_ENERGY] = energy;
    return includeEnergy ? energy : 0.0;
}

double ATMMetaForceImpl::calcForcesAndEnergy(OpenMM::ContextImpl &context, bool includeForces, bool includeEnergy, int groups) {
    if (kernel) return calcWithPlatformKernel(context, includeForces, includeEnergy, groups);
    if ((groups & (1 << owner.getForceGroup())) == 0) return 0.0;
    if (!handle) createBackend(context);
    const int n = owner.getNumParticles();
    const std::vector<OpenMM::Vec3> &pos = context.positionsRef();
    // the nine global parameters, by name (ref: CommonATMMetaForceKernels.cpp:165-181 reads them from the context)
    double p[ATM_NUM_PARAMS] = {context.getParameter(ATMMetaForce::Lambda1()), context.getParameter(ATMMetaForce::Lambda2()),
                                context.getParameter(ATMMetaForce::Alpha()),   context.getParameter(ATMMetaForce::U0()),
                                context.getParameter(ATMMetaForce::W0()),      context.getParameter(ATMMetaForce::Umax()),
                                context.getParameter(ATMMetaForce::Ubcore()),  context.getParameter(ATMMetaForce::Acore()),
                                context.getParameter(ATMMetaForce::Direction())};
    check(atm_set_parameters(handle, 0, p), "ATMMetaForce: parameters");
    bool rebuild = refRebuild.empty();
    if (displacementsDirty) {
        displacements = owner.getDisplacementArray();
        check(atm_set_displacements(handle, nullptr, displacements.data(), stream), "ATMMetaForce: uploading the displacement table");
        displacementsDirty = false;
        rebuild = true;
    }
    if (boxVersionSeen != context.getBoxVersion()) {
        OpenMM::Vec3 a, b, c;
        context.getPeriodicBoxVectors(a, b, c);
        const double box[9] = {a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]};
        check(atm_set_box(handle, -1, box), "ATMMetaForce: periodic box");
        if (reciprocalOn) {   // the mesh follows the box
            OpenMM::Vec3 bv[3] = {a, b, c};
            int grid[3];
            OpenMM::pmeGridDimensions(std::sqrt(-std::log(2.0 * nonbonded->getEwaldErrorTolerance())) / nonbonded->getCutoffDistance(),
                                      nonbonded->getEwaldErrorTolerance(), bv, grid);
            check(atm_stream_synchronize(stream), "ATMMetaForce: waiting before the mesh change");
            check(atm_pme_setup(handle, grid[0], grid[1], grid[2], 5), "ATMMetaForce: PME mesh");
        }
        boxVersionSeen = context.getBoxVersion();
        rebuild = true;
    }
    // pair lists: rebuild / prune when some atom may have moved by more than half the respective skin
    int maintenance = 0;
    if (rebuild || maxMove(pos, refRebuild) > 0.5 * skinOuter) {
        maintenance = 2;
        refRebuild = pos;
        refPrune = pos;
    } else if (maxMove(pos, refPrune) > 0.5 * skin) {
        maintenance = 1;
        refPrune = pos;
    }
    for (int i = 0; i < n; i++) {
        posqHost[4 * i] = (float)pos[i][0];
        posqHost[4 * i + 1] = (float)pos[i][1];
        posqHost[4 * i + 2] = (float)pos[i][2];
    }
    atm_host_io io;
    io.posq_host = posqHost;
    io.force_host = forceHost;
    io.energies_host = energyHost;
    io.include_energy = 1;   // the reference always evaluates the inner energies (do_energy = true, :104)
    io.force_format = ATM_FORCE_I64;
    io.posq_format = ATM_POSQ_F4;
    io.reserved = 0;
    io.force_state1_ext_host = io.force_state2_ext_host = nullptr;
    io.energy_ext_host = nullptr;
    if (!hostForces.empty()) {
        // the other variable-group Forces at the state-1 (x) and state-2 (x + d) coordinates: what the reference's inner
        // contexts add to State1Energy / State2Energy and to their force buffers (ref: :113-116)
        OpenMM::Vec3 box[3];
        context.getPeriodicBoxVectors(box[0], box[1], box[2]);
        std::vector<OpenMM::Vec3> statePos(pos), f(n);
        for (int s = 0; s < 2; s++) {
            if (s == 1)
                for (int i = 0; i < n; i++)
                    for (int c = 0; c < 3; c++) statePos[i][c] = pos[i][c] + displacements[3 * (size_t)i + c];
            std::fill(f.begin(), f.end(), OpenMM::Vec3());
            double e = 0.0;
            for (const OpenMM::HostEvaluatedForce *hf : hostForces) e += hf->evaluate(statePos, box, &f);
            energyExtHost[s] = e;
            for (int i = 0; i < n; i++)
                for (int c = 0; c < 3; c++) extHost[s][(size_t)c * paddedNumAtoms + i] = (int64_t)std::llrint(f[i][c] * 4294967296.0);
        }
        io.force_state1_ext_host = extHost[0];
        io.force_state2_ext_host = extHost[1];
        io.energy_ext_host = energyExtHost;
    }
    check(atm_host_pipeline_step(pipeline, &io, maintenance, stream), "ATMMetaForce: evaluating the alchemical force");
    check(atm_stream_synchronize(stream), "ATMMetaForce: waiting for the step");
    if (maintenance == 2) {
        // the rebuild inside that step was asynchronous: if a pair list outgrew its capacity the step returned NaN;
        // repeat it (the rebuild now reallocates and verifies synchronously) before anything reaches the integrator
        const int rc = atm_host_pipeline_check(pipeline);
        if (rc == ATM_ERR_STATE) {
            check(atm_host_pipeline_step(pipeline, &io, 2, stream), "ATMMetaForce: repeating the step after a pair-list capacity change");
            check(atm_stream_synchronize(stream), "ATMMetaForce: waiting for the step");
            check(atm_host_pipeline_check(pipeline), "ATMMetaForce: pair lists");
        } else {
            check(rc, "ATMMetaForce: pair lists");
        }
    }
    energyRecord.assign(energyHost, energyHost + ATM_NUM_ENERGY_SLOTS);
    PerturbationEnergy = energyRecord[ATM_E_USC];
    if (includeForces) {
        std::vector<OpenMM::Vec3> &f = context.getForces();
        const double inv = 1.0 / 4294967296.0;   // 2^32 fixed point of the long force buffers
        for (int i = 0; i < n; i++)
            for (int c = 0; c < 3; c++) f[i][c] += (double)forceHost[(size_t)c * paddedNumAtoms + i] * inv;
    }
    return includeEnergy ? energyRecord[ATM_E_ENERGY] : 0.0;
}

// ---- the three members of ATMMetaForce that need the Impl (ref: openmmapi/src/ATMMetaForce.cpp:34-44)

OpenMM::ForceImpl *ATMMetaForce::createImpl() const { return new ATMMetaForceImpl(*this); }

void ATMMetaForce::updateParametersInContext(OpenMM::Context &context) {
    dynamic_cast<ATMMetaForceImpl &>(context.getForceImpl(*this)).updateParametersInContext(context.getImpl());
}

double ATMMetaForce::getPerturbationEnergy(const OpenMM::Context &context) const {
    return dynamic_cast<const ATMMetaForceImpl &>(context.getForceImpl(*this)).getPerturbationEnergy();
}

#endif  // !ATM_HAVE_OPENMM
