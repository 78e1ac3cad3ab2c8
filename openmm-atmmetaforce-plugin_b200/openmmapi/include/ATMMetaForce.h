// ATMMetaForce.h -- the public Force class of the plugin, same namespace / name / members as the reference
// (ref: openmmapi/include/ATMMetaForce.h:54-315), written from the API description in SURVEY.md section 8b.
//
// An ATMMetaForce defines the Alchemical Transfer potential
//     E = e0 + W(u_sc),   u = +-(U2 - U1),   U2 = U(x + d)
// for the forces of the System that live in the "variable force groups": U1 is their energy at the current
// coordinates x and U2 at the coordinates displaced per particle by (dx,dy,dz).  The nine scalars below become
// Context global parameters under the names returned by Lambda1() ... Direction().
#ifndef OPENMM_ATMMETAFORCE_H_
#define OPENMM_ATMMETAFORCE_H_

#ifdef ATM_HAVE_OPENMM
#include "openmm/Force.h"
#else
#include "openmm_standin.h"
#endif
#include <string>
#include <vector>
#include "ATMMetaForceVersion.h"

namespace ATMMetaForcePlugin {

class ATMMetaForce : public OpenMM::Force {
public:
    /**
     * @param lambda1, lambda2   softplus slopes (dimensionless)
     * @param alpha              softplus sharpness, (kJ/mol)^-1
     * @param u0, w0             softplus offset and constant, kJ/mol
     * @param umax, ubcore, acore soft-core ceiling, onset (kJ/mol) and exponent
     * @param direction          +1: x is the reference state, x+d the displaced one; -1: roles swapped
     * @param VariableForceGroups force groups re-evaluated at the displaced coordinates
     */
    ATMMetaForce(double lambda1, double lambda2, double alpha, double u0, double w0, double umax, double ubcore,
                 double acore, double direction, const std::vector<int> &VariableForceGroups)
        : defaultLambda1(lambda1), defaultLambda2(lambda2), defaultAlpha(alpha), defaultU0(u0), defaultW0(w0),
          defaultUmax(umax), defaultUbcore(ubcore), defaultAcore(acore), defaultDirection(direction),
          VariableForceGroups(VariableForceGroups) {}

    int getNumParticles() const { return (int)particles.size(); }
    /** Appends a particle with its displacement (nm); returns its index in the force. */
    int addParticle(int particle, double dx, double dy, double dz);
    void getParticleParameters(int index, int &particle, double &dx, double &dy, double &dz) const;
    void setParticleParameters(int index, int particle, double dx, double dy, double dz);
    bool usesPeriodicBoundaryConditions() const { return false; }

    // names of the Context global parameters
    static const std::string &Lambda1() { static const std::string k = "ATMLambda1"; return k; }
    static const std::string &Lambda2() { static const std::string k = "ATMLambda2"; return k; }
    static const std::string &Alpha() { static const std::string k = "ATMAlpha"; return k; }
    static const std::string &U0() { static const std::string k = "ATMU0"; return k; }
    static const std::string &W0() { static const std::string k = "ATMW0"; return k; }
    static const std::string &Umax() { static const std::string k = "ATMUmax"; return k; }
    static const std::string &Ubcore() { static const std::string k = "ATMUbcore"; return k; }
    static const std::string &Acore() { static const std::string k = "ATMAcore"; return k; }
    static const std::string &Direction() { static const std::string k = "ATMDirection"; return k; }
    static const std::string &Version() { static const std::string v = ATMMETAFORCE_VERSION; return v; }

    double getDefaultLambda1() const { return defaultLambda1; }
    double getDefaultLambda2() const { return defaultLambda2; }
    double getDefaultAlpha() const { return defaultAlpha; }
    double getDefaultU0() const { return defaultU0; }
    double getDefaultW0() const { return defaultW0; }
    double getDefaultUmax() const { return defaultUmax; }
    double getDefaultUbcore() const { return defaultUbcore; }
    double getDefaultAcore() const { return defaultAcore; }
    double getDefaultDirection() const { return defaultDirection; }
    const std::vector<int> &getVariableForceGroups() const { return VariableForceGroups; }

    /** Pushes changed per-particle displacements into an existing Context (the number of particles cannot change).
     *  ref: ATMMetaForce::updateParametersInContext (openmmapi/include/ATMMetaForce.h:130, src/ATMMetaForce.cpp:38-40). */
    void updateParametersInContext(OpenMM::Context &context);
    /** Soft-core perturbation energy u_sc (kJ/mol) of the Context's last evaluation.
     *  ref: ATMMetaForce::getPerturbationEnergy (ATMMetaForce.h:145, src/ATMMetaForce.cpp:42-44). */
    double getPerturbationEnergy(const OpenMM::Context &context) const;

    /** The nine defaults in atm_b200.h parameter order (lambda1 ... direction). */
    void getDefaultParameters(double p[9]) const;
    /** Displacements as a dense [numParticles][3] array indexed by FORCE ENTRY (entry i == atom i; the 'particle'
     *  field is stored but, exactly like the reference's GPU platforms, not used for indexing). */
    std::vector<double> getDisplacementArray() const;

#ifdef ATM_HAVE_OPENMM
protected:
#endif
    /** ref: ATMMetaForce::createImpl (ATMMetaForce.h:300, src/ATMMetaForce.cpp:34-36). */
    OpenMM::ForceImpl *createImpl() const;

private:
    struct ParticleInfo {
        int particle;
        double dx, dy, dz;
    };
    std::vector<ParticleInfo> particles;
    double defaultLambda1, defaultLambda2, defaultAlpha, defaultU0, defaultW0;
    double defaultUmax, defaultUbcore, defaultAcore;
    double defaultDirection;
    std::vector<int> VariableForceGroups;
};

}  // namespace ATMMetaForcePlugin

#endif
