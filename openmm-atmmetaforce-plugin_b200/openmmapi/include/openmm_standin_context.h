// openmm_standin_context.h -- System / NonbondedForce / Context for builds WITHOUT OpenMM, just enough for
// ATMMetaForceImpl to be driven exactly the way OpenMM drives a ForceImpl (initialize once, calcForcesAndEnergy per
// evaluation, global parameters by name).  Method names and argument orders follow OpenMM's documented public API
// (System::addParticle/addForce, NonbondedForce::addParticle/addException/setCutoffDistance/setEwaldErrorTolerance,
// Context::setPositions/setPeriodicBoxVectors/setParameter/getParameter); nothing here is copied from OpenMM.
// With -DATM_HAVE_OPENMM this file is not used.
#ifndef ATM_OPENMM_STANDIN_CONTEXT_H_
#define ATM_OPENMM_STANDIN_CONTEXT_H_

#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "openmm_standin.h"

namespace OpenMM {

struct Vec3 {
    double v[3];
    Vec3() : v{0.0, 0.0, 0.0} {}
    Vec3(double x, double y, double z) : v{x, y, z} {}
    double operator[](int i) const { return v[i]; }
    double &operator[](int i) { return v[i]; }
};

class ContextImpl;

/** What OpenMM calls a ForceImpl: the per-Context implementation object of a Force. */
class ForceImpl {
public:
    virtual ~ForceImpl() {}
    virtual void initialize(ContextImpl &context) = 0;
    virtual double calcForcesAndEnergy(ContextImpl &context, bool includeForces, bool includeEnergy, int groups) = 0;
    virtual std::map<std::string, double> getDefaultParameters() = 0;
    virtual std::vector<std::string> getKernelNames() = 0;
};

/** The subset of OpenMM's NonbondedForce the Tier-2 path evaluates: charges, Lennard-Jones sigma / epsilon
 *  (Lorentz-Berthelot), exceptions (each one excludes its pair from the regular sum; a non-zero chargeProd or epsilon
 *  makes it a scaled 1-4 interaction), periodic cutoff with Ewald real space. */
class NonbondedForce : public Force {
public:
    enum NonbondedMethod { NoCutoff = 0, CutoffNonPeriodic = 1, CutoffPeriodic = 2, Ewald = 3, PME = 4 };
    NonbondedForce() : method(PME), cutoff(1.0), ewaldTolerance(5e-4) {}
    int addParticle(double charge, double sigma, double epsilon) {
        charges.push_back(charge); sigmas.push_back(sigma); epsilons.push_back(epsilon);
        return (int)charges.size() - 1;
    }
    int getNumParticles() const { return (int)charges.size(); }
    void getParticleParameters(int index, double &charge, double &sigma, double &epsilon) const {
        ASSERT_VALID_INDEX(index, charges);
        charge = charges[index]; sigma = sigmas[index]; epsilon = epsilons[index];
    }
    int addException(int particle1, int particle2, double chargeProd, double sigma, double epsilon) {
        exceptionPairs.push_back(particle1); exceptionPairs.push_back(particle2);
        exceptionParams.push_back(chargeProd); exceptionParams.push_back(sigma); exceptionParams.push_back(epsilon);
        return (int)exceptionPairs.size() / 2 - 1;
    }
    int getNumExceptions() const { return (int)exceptionPairs.size() / 2; }
    void getExceptionParameters(int index, int &particle1, int &particle2, double &chargeProd, double &sigma, double &epsilon) const {
        if (index < 0 || index >= getNumExceptions()) throw OpenMMException("Assertion failure: Index out of range");
        particle1 = exceptionPairs[2 * index]; particle2 = exceptionPairs[2 * index + 1];
        chargeProd = exceptionParams[3 * index]; sigma = exceptionParams[3 * index + 1]; epsilon = exceptionParams[3 * index + 2];
    }
    NonbondedMethod getNonbondedMethod() const { return method; }
    void setNonbondedMethod(NonbondedMethod m) { method = m; }
    double getCutoffDistance() const { return cutoff; }
    void setCutoffDistance(double distance) { cutoff = distance; }
    double getEwaldErrorTolerance() const { return ewaldTolerance; }
    void setEwaldErrorTolerance(double tol) { ewaldTolerance = tol; }
    bool usesPeriodicBoundaryConditions() const override { return method >= CutoffPeriodic; }

private:
    NonbondedMethod method;
    double cutoff, ewaldTolerance;
    std::vector<double> charges, sigmas, epsilons, exceptionParams;
    std::vector<int> exceptionPairs;
};

/** Particles (masses), forces (owned), default box. */
class System {
public:
    int addParticle(double mass) { masses.push_back(mass); return (int)masses.size() - 1; }
    int getNumParticles() const { return (int)masses.size(); }
    double getParticleMass(int index) const { ASSERT_VALID_INDEX(index, masses); return masses[index]; }
    /** Takes ownership of the force, as OpenMM::System does. */
    int addForce(Force *force) { forces.emplace_back(force); return (int)forces.size() - 1; }
    int getNumForces() const { return (int)forces.size(); }
    const Force &getForce(int index) const { ASSERT_VALID_INDEX(index, forces); return *forces[index]; }
    Force &getForce(int index) { ASSERT_VALID_INDEX(index, forces); return *forces[index]; }
    void setDefaultPeriodicBoxVectors(const Vec3 &a, const Vec3 &b, const Vec3 &c) { box[0] = a; box[1] = b; box[2] = c; }
    void getDefaultPeriodicBoxVectors(Vec3 &a, Vec3 &b, Vec3 &c) const { a = box[0]; b = box[1]; c = box[2]; }

private:
    std::vector<double> masses;
    std::vector<std::unique_ptr<Force>> forces;
    Vec3 box[3];
};

/** The state a ForceImpl sees. */
class ContextImpl {
public:
    explicit ContextImpl(const System &system) : system(system), positionsSet(false), boxVersion(0) {
        system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
    }
    const System &getSystem() const { return system; }
    double getParameter(const std::string &name) const {
        auto it = parameters.find(name);
        if (it == parameters.end()) throw OpenMMException("Called getParameter() with invalid parameter name: " + name);
        return it->second;
    }
    void setParameter(const std::string &name, double value) {
        auto it = parameters.find(name);
        if (it == parameters.end()) throw OpenMMException("Called setParameter() with invalid parameter name: " + name);
        it->second = value;
    }
    const std::map<std::string, double> &getParameters() const { return parameters; }
    void getPositions(std::vector<Vec3> &out) const { out = positions; }
    const std::vector<Vec3> &positionsRef() const { return positions; }
    bool hasPositions() const { return positionsSet; }
    void getPeriodicBoxVectors(Vec3 &a, Vec3 &b, Vec3 &c) const { a = box[0]; b = box[1]; c = box[2]; }
    unsigned long getBoxVersion() const { return boxVersion; }
    /** Force accumulator of the current evaluation, kJ/mol/nm, one Vec3 per particle. */
    std::vector<Vec3> &getForces() { return forces; }

private:
    friend class Context;
    const System &system;
    std::map<std::string, double> parameters;
    std::vector<Vec3> positions, forces;
    Vec3 box[3];
    bool positionsSet;
    unsigned long boxVersion;
};

/** Owns one ForceImpl per Force that has one; evaluates force groups the way Context::getState does. */
class Context {
public:
    explicit Context(const System &system) : impl(system) {
        for (int i = 0; i < system.getNumForces(); i++) {
            ForceImpl *fi = system.getForce(i).createImpl();
            if (!fi) continue;
            forceImpls.emplace_back(&system.getForce(i), std::unique_ptr<ForceImpl>(fi));
            for (const auto &kv : fi->getDefaultParameters()) impl.parameters[kv.first] = kv.second;
        }
        for (auto &fi : forceImpls) fi.second->initialize(impl);
    }
    const System &getSystem() const { return impl.getSystem(); }
    void setPositions(const std::vector<Vec3> &positions) {
        if ((int)positions.size() != impl.getSystem().getNumParticles())
            throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
        impl.positions = positions;
        impl.positionsSet = true;
    }
    void setPeriodicBoxVectors(const Vec3 &a, const Vec3 &b, const Vec3 &c) {
        impl.box[0] = a; impl.box[1] = b; impl.box[2] = c;
        impl.boxVersion++;
    }
    void setParameter(const std::string &name, double value) { impl.setParameter(name, value); }
    double getParameter(const std::string &name) const { return impl.getParameter(name); }
    const std::map<std::string, double> &getParameters() const { return impl.getParameters(); }
    /** Energy (kJ/mol) of the force groups in the bit mask `groups`; forces are left in getForces(). */
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = -1) {
        if (!impl.positionsSet) throw OpenMMException("Particle positions have not been set");
        impl.forces.assign(impl.getSystem().getNumParticles(), Vec3());
        double energy = 0.0;
        for (auto &fi : forceImpls) energy += fi.second->calcForcesAndEnergy(impl, includeForces, includeEnergy, groups);
        return energy;
    }
    const std::vector<Vec3> &getForces() const { return impl.forces; }
    ContextImpl &getImpl() { return impl; }
    const ContextImpl &getImpl() const { return impl; }
    ForceImpl &getForceImpl(const Force &force) const {
        for (auto &fi : forceImpls)
            if (fi.first == &force) return *fi.second;
        throw OpenMMException("The Force is not part of this Context's System");
    }

private:
    ContextImpl impl;
    std::vector<std::pair<const Force *, std::unique_ptr<ForceImpl>>> forceImpls;
};

}  // namespace OpenMM

#endif
