#include "ATMMetaForceB200Kernel.h"

using namespace ATMMetaForcePlugin;
using OpenMM::OpenMMException;

static void check(int rc, const char *what) {
    if (rc != ATM_OK) throw OpenMMException(std::string(what) + ": " + atm_last_error());
}

ATMMetaForceB200Kernel::~ATMMetaForceB200Kernel() {
    if (handle) atm_destroy(handle);
}

void ATMMetaForceB200Kernel::initialize(const ATMMetaForce &force, int paddedNumAtoms, atm_precision precision,
                                        const std::vector<int> &index, int device) {
    if (handle) {
        atm_destroy(handle);
        handle = nullptr;
    }
    numParticles = force.getNumParticles();
    atm_config cfg;
    cfg.num_particles = numParticles;
    cfg.padded_num_particles = paddedNumAtoms;
    cfg.precision = precision;
    cfg.num_replicas = 1;
    cfg.device = device;
    check(atm_create(&cfg, &handle), "ATMMetaForce: creating the Blackwell back-end");
    atomsReordered(force, index, nullptr);
    double p[9];
    force.getDefaultParameters(p);
    check(atm_set_parameters(handle, 0, p), "ATMMetaForce: default parameters");
}

void ATMMetaForceB200Kernel::atomsReordered(const ATMMetaForce &force, const std::vector<int> &index, void *stream) {
    if (force.getNumParticles() != numParticles)
        throw OpenMMException("copyParametersToContext: The number of ATMMetaForce particles has changed");
    atomIndex = index;
    if (!atomIndex.empty() && (int)atomIndex.size() != numParticles)
        throw OpenMMException("ATMMetaForce: atom index has the wrong length");
    std::vector<double> d = force.getDisplacementArray();
    check(atm_set_displacements(handle, atomIndex.empty() ? nullptr : atomIndex.data(), d.data(), stream),
          "ATMMetaForce: uploading the displacement table");
}

void ATMMetaForceB200Kernel::copyParametersToContext(const ATMMetaForce &force, void *stream) {
    atomsReordered(force, atomIndex, stream);
}

void ATMMetaForceB200Kernel::copyState(const void *posq, const void *posqCorrection, void *posq1, void *posq1Correction,
                                       void *posq2, void *posq2Correction, void *stream) {
    check(atm_copy_state(handle, posq, posqCorrection, posq1, posq1Correction, posq2, posq2Correction, stream),
          "ATMMetaForce: copyState");
}

double ATMMetaForceB200Kernel::execute(const std::map<std::string, double> &parameters, long long *force,
                                       const long long *forceState1, const long long *forceState2, double State1Energy,
                                       double State2Energy, bool /*includeForces*/, bool includeEnergy, void *stream) {
    // the reference ignores includeForces as well: the inner forces were already computed
    const std::string *names[9] = {&ATMMetaForce::Lambda1(), &ATMMetaForce::Lambda2(), &ATMMetaForce::Alpha(),
                                   &ATMMetaForce::U0(), &ATMMetaForce::W0(), &ATMMetaForce::Umax(),
                                   &ATMMetaForce::Ubcore(), &ATMMetaForce::Acore(), &ATMMetaForce::Direction()};
    double p[9];
    for (int k = 0; k < 9; k++) {
        auto it = parameters.find(*names[k]);
        if (it == parameters.end()) throw OpenMMException("Called getParameter() with invalid parameter name: " + *names[k]);
        p[k] = it->second;
    }
    check(atm_set_parameters(handle, 0, p), "ATMMetaForce: parameters");
    double energy = 0.0;
    check(atm_execute(handle, 0, State1Energy, State2Energy, (int64_t *)force, (const int64_t *)forceState1,
                      (const int64_t *)forceState2, includeEnergy ? 1 : 0, &energy, stream),
          "ATMMetaForce: execute");
    check(atm_get_perturbation_energy(handle, 0, &perturbationEnergy), "ATMMetaForce: perturbation energy");
    return energy;
}

std::map<std::string, double> ATMMetaForceB200Kernel::getDefaultParameters(const ATMMetaForce &force) {
    std::map<std::string, double> m;
    m[ATMMetaForce::Lambda1()] = force.getDefaultLambda1();
    m[ATMMetaForce::Lambda2()] = force.getDefaultLambda2();
    m[ATMMetaForce::Alpha()] = force.getDefaultAlpha();
    m[ATMMetaForce::U0()] = force.getDefaultU0();
    m[ATMMetaForce::W0()] = force.getDefaultW0();
    m[ATMMetaForce::Umax()] = force.getDefaultUmax();
    m[ATMMetaForce::Ubcore()] = force.getDefaultUbcore();
    m[ATMMetaForce::Acore()] = force.getDefaultAcore();
    m[ATMMetaForce::Direction()] = force.getDefaultDirection();
    return m;
}

int ATMMetaForcePlugin::variableForceGroupsMask(const ATMMetaForce &force) {
    int mask = 0;
    for (int g : force.getVariableForceGroups()) {
        if (g == force.getForceGroup())
            throw OpenMMException("The ATM Meta Force group cannot be one of the variable force groups.");
        if (g < 0 || g > 31) throw OpenMMException("Force group must be between 0 and 31");
        mask |= 1 << g;  // the reference sums with += (a duplicated id corrupts its mask); OR is idempotent
    }
    return mask;
}
