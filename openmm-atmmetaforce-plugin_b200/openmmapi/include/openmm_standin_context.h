// openmm_standin_context.h -- System / NonbondedForce / Context for builds WITHOUT OpenMM, just enough for
// ATMMetaForceImpl to be driven exactly the way OpenMM drives a ForceImpl (initialize once, calcForcesAndEnergy per
// evaluation, global parameters by name).  Method names and argument orders follow OpenMM's documented public API
// (System::addParticle/addForce, NonbondedForce::addParticle/addException/setCutoffDistance/setEwaldErrorTolerance,
// Context::setPositions/setPeriodicBoxVectors/setParameter/getParameter, Platform::registerKernelFactory/createKernel/
// getPlatformByName/registerPlatform, KernelImpl / Kernel::getAs / KernelFactory::createKernelImpl,
// ContextImpl::getPlatform/getPlatformData/createLinkedContext/calcForcesAndEnergy); nothing here is copied from OpenMM.
// Two platforms exist in this build: the built-in "HostB200" (no kernels: ATMMetaForceImpl evaluates the fused Tier-2
// path from host positions) and the stand-in "CUDA" platform of openmm_standin_cuda.h (device-resident posq / long
// force buffers; ATMMetaForceImpl then runs the reference's own orchestration -- two linked inner contexts and the
// CalcATMMetaForceKernel seam registered by libATMMetaForcePluginCUDA.so).
// With -DATM_HAVE_OPENMM this file is not used.
#ifndef ATM_OPENMM_STANDIN_CONTEXT_H_
#define ATM_OPENMM_STANDIN_CONTEXT_H_

#include <cmath>
#include <functional>
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
    NonbondedForce() : method(PME), cutoff(1.0), ewaldTolerance(5e-4), recipGroup(-1), dispersionCorrection(true) {}
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
    /** -1 = reciprocal space belongs to the force's own group (what ATMMetaForceImpl::copysystem sets on its clones). */
    int getReciprocalSpaceForceGroup() const { return recipGroup; }
    void setReciprocalSpaceForceGroup(int group) { recipGroup = group; }
    bool getUseDispersionCorrection() const { return dispersionCorrection; }
    void setUseDispersionCorrection(bool on) { dispersionCorrection = on; }
    Force *clone() const override { return new NonbondedForce(*this); }
    /** On a platform with a "CalcNonbondedForce" kernel this is an ordinary Force with an Impl of its own; on the
     *  kernel-less host platform it is only data for ATMMetaForceImpl (the Impl is inert). */
    ForceImpl *createImpl() const override;

private:
    NonbondedMethod method;
    double cutoff, ewaldTolerance;
    int recipGroup;
    bool dispersionCorrection;
    std::vector<double> charges, sigmas, epsilons, exceptionParams;
    std::vector<int> exceptionPairs;
};

/** A Force this OpenMM-free build evaluates on the HOST: the stand-in for OpenMM's bonded / custom force kernels, so that
 *  the variable force groups can hold something besides the NonbondedForce (the reference's inner contexts evaluate
 *  whatever Force sits in a variable group, ref: openmmapi/src/ATMMetaForceImpl.cpp:51-65,113-116).  Its Impl works on
 *  every platform (positions come from / forces go to the platform's hooks); the fused ATMMetaForceImpl evaluates it
 *  at the state-1 and state-2 coordinates and feeds the results to the back-end as the external per-state terms. */
class HostEvaluatedForce : public Force {
public:
    /** Energy (kJ/mol) at `positions`; when forces != NULL the forces (kJ/mol/nm) are ADDED to *forces. */
    virtual double evaluate(const std::vector<Vec3> &positions, const Vec3 box[3], std::vector<Vec3> *forces) const = 0;
    ForceImpl *createImpl() const override;
};

/** OpenMM's HarmonicBondForce: E = 1/2 k (r - r0)^2 per bond. */
class HarmonicBondForce : public HostEvaluatedForce {
public:
    HarmonicBondForce() : periodic(false) {}
    int addBond(int particle1, int particle2, double length, double k) {
        bonds.push_back({particle1, particle2, length, k});
        return (int)bonds.size() - 1;
    }
    int getNumBonds() const { return (int)bonds.size(); }
    void getBondParameters(int index, int &particle1, int &particle2, double &length, double &k) const {
        ASSERT_VALID_INDEX(index, bonds);
        particle1 = bonds[index].p1; particle2 = bonds[index].p2; length = bonds[index].length; k = bonds[index].k;
    }
    void setBondParameters(int index, int particle1, int particle2, double length, double k) {
        ASSERT_VALID_INDEX(index, bonds);
        bonds[index] = {particle1, particle2, length, k};
    }
    void setUsesPeriodicBoundaryConditions(bool on) { periodic = on; }
    bool usesPeriodicBoundaryConditions() const override { return periodic; }
    Force *clone() const override { return new HarmonicBondForce(*this); }
    double evaluate(const std::vector<Vec3> &positions, const Vec3 box[3], std::vector<Vec3> *forces) const override;

private:
    struct Bond { int p1, p2; double length, k; };
    std::vector<Bond> bonds;
    bool periodic;
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
    int addConstraint(int particle1, int particle2, double distance) {
        constraints.push_back({particle1, particle2, distance});
        return (int)constraints.size() - 1;
    }
    int getNumConstraints() const { return (int)constraints.size(); }
    void getConstraintParameters(int index, int &particle1, int &particle2, double &distance) const {
        ASSERT_VALID_INDEX(index, constraints);
        particle1 = constraints[index].p1; particle2 = constraints[index].p2; distance = constraints[index].d;
    }
    void setDefaultPeriodicBoxVectors(const Vec3 &a, const Vec3 &b, const Vec3 &c) { box[0] = a; box[1] = b; box[2] = c; }
    void getDefaultPeriodicBoxVectors(Vec3 &a, Vec3 &b, Vec3 &c) const { a = box[0]; b = box[1]; c = box[2]; }

private:
    struct Constraint { int p1, p2; double d; };
    std::vector<double> masses;
    std::vector<Constraint> constraints;
    std::vector<std::unique_ptr<Force>> forces;
    Vec3 box[3];
};

/** Step size only: the inner contexts of ATMMetaForceImpl are never integrated (ref: ATMMetaForceImpl.cpp:27). */
class Integrator {
public:
    explicit Integrator(double stepSize) : stepSize(stepSize) {}
    virtual ~Integrator() {}
    double getStepSize() const { return stepSize; }

private:
    double stepSize;
};
class VerletIntegrator : public Integrator {
public:
    explicit VerletIntegrator(double stepSize) : Integrator(stepSize) {}
};

class Platform;
class Context;

/** A platform-specific implementation of a named computation (OpenMM::KernelImpl). */
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform &platform) : name(std::move(name)), platform(&platform) {}
    virtual ~KernelImpl() {}
    const std::string &getName() const { return name; }
    const Platform &getPlatform() const { return *platform; }

private:
    std::string name;
    const Platform *platform;
};

/** Shared handle to a KernelImpl (OpenMM::Kernel). */
class Kernel {
public:
    Kernel() {}
    explicit Kernel(KernelImpl *impl) : impl(impl) {}
    const std::string &getName() const { return impl->getName(); }
    KernelImpl &getImpl() { return *impl; }
    template <class T>
    T &getAs() { return dynamic_cast<T &>(*impl); }
    template <class T>
    const T &getAs() const { return dynamic_cast<const T &>(*impl); }
    explicit operator bool() const { return (bool)impl; }

private:
    std::shared_ptr<KernelImpl> impl;
};

class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl *createKernelImpl(std::string name, const Platform &platform, ContextImpl &context) const = 0;
};

/** Kernel factories by name plus the global platform registry (OpenMM::Platform).  The virtual hooks at the end condense
 *  what OpenMM does through its "UpdateStateData" / "CalcForcesAndEnergy" kernels: where positions and forces live. */
class Platform {
public:
    virtual ~Platform() {}
    virtual const std::string &getName() const = 0;
    /** Takes ownership of the factory (a factory registered for several names is shared). */
    void registerKernelFactory(const std::string &name, KernelFactory *factory) {
        for (auto &kv : factories)
            if (kv.second.get() == factory) { factories[name] = kv.second; return; }
        factories[name] = std::shared_ptr<KernelFactory>(factory);
    }
    bool supportsKernels(const std::vector<std::string> &names) const {
        for (const auto &n : names)
            if (!factories.count(n)) return false;
        return true;
    }
    Kernel createKernel(const std::string &name, ContextImpl &context) const {
        auto it = factories.find(name);
        if (it == factories.end()) throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
        return Kernel(it->second->createKernelImpl(name, *this, context));
    }
    static void registerPlatform(Platform *platform) { registry().emplace_back(platform); }
    static int getNumPlatforms() { return (int)registry().size(); }
    static Platform &getPlatform(int index) { ASSERT_VALID_INDEX(index, registry()); return *registry()[index]; }
    static Platform &getPlatformByName(const std::string &name) {
        for (auto &p : registry())
            if (p->getName() == name) return *p;
        throw OpenMMException("There is no registered Platform called \"" + name + "\"");
    }
    /** dlopen()s a plugin and calls its registerPlatforms() and registerKernelFactories() (OpenMM::Platform::loadPluginLibrary). */
    static void loadPluginLibrary(const std::string &file);

    // ---- where a Context's state lives on this platform
    virtual void contextCreated(ContextImpl &context, const std::map<std::string, std::string> &properties) const;
    virtual void linkedContextCreated(ContextImpl &context, ContextImpl &original) const;
    virtual void contextDestroyed(ContextImpl &context) const;
    virtual void setPositions(ContextImpl &context, const std::vector<Vec3> &positions) const;
    virtual void getPositions(const ContextImpl &context, std::vector<Vec3> &positions) const;
    virtual void beginComputation(ContextImpl &context, bool includeForces, bool includeEnergy, int groups) const;
    virtual double finishComputation(ContextImpl &context, bool includeForces, bool includeEnergy, int groups) const;
    virtual void getForces(ContextImpl &context, std::vector<Vec3> &forces) const;
    /** Adds host-evaluated forces (atom order, kJ/mol/nm) to the forces of the evaluation in progress. */
    virtual void addForces(ContextImpl &context, const std::vector<Vec3> &forces) const;

private:
    /** ONE registry per process: defined in openmm_standin_context.cpp, i.e. inside libOpenMMStandin.so, which the API
     *  library, every plugin library and the host all link (as they link libOpenMM.so in a real installation). */
    static std::vector<std::unique_ptr<Platform>> &registry();
    std::map<std::string, std::shared_ptr<KernelFactory>> factories;
};

/** Kernel seam of the stand-in NonbondedForce (OpenMM declares CalcNonbondedForceKernel in kernels.h). */
class CalcNonbondedForceKernel : public KernelImpl {
public:
    static std::string Name() { return "CalcNonbondedForce"; }
    CalcNonbondedForceKernel(std::string name, const Platform &platform) : KernelImpl(name, platform) {}
    virtual void initialize(const System &system, const NonbondedForce &force) = 0;
    virtual double execute(ContextImpl &context, bool includeForces, bool includeEnergy, bool includeDirect, bool includeReciprocal) = 0;
};

/** OpenMM's rule for the PME mesh: ceil(2 alpha L / (3 tol^(1/5))) per axis, rounded up to a 2/3/5/7-smooth size. */
void pmeGridDimensions(double alpha, double tolerance, const Vec3 box[3], int grid[3]);

/** The kernel-less host platform: positions and forces are the host vectors of the ContextImpl. */
class HostPlatform : public Platform {
public:
    const std::string &getName() const override {
        static const std::string name = "HostB200";
        return name;
    }
    static HostPlatform &instance();
};

/** The state a ForceImpl sees (OpenMM::ContextImpl): system, platform, parameters, time, box, force impls. */
class ContextImpl {
public:
    ContextImpl(Context &owner, const System &system, Platform &platform, const std::map<std::string, std::string> &properties,
                ContextImpl *originalContext);
    ~ContextImpl();
    ContextImpl(const ContextImpl &) = delete;
    ContextImpl &operator=(const ContextImpl &) = delete;
    Context &getOwner() { return owner; }
    const System &getSystem() const { return system; }
    Platform &getPlatform() { return *platform; }
    const Platform &getPlatform() const { return *platform; }
    void *getPlatformData() { return platformData; }
    const void *getPlatformData() const { return platformData; }
    void setPlatformData(void *data) { platformData = data; }
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
    double getTime() const { return time; }
    void setTime(double t) { time = t; }
    void setPositions(const std::vector<Vec3> &p) {
        if ((int)p.size() != system.getNumParticles())
            throw OpenMMException("Called setPositions() on a Context with the wrong number of positions");
        positions = p;
        positionsSet = true;
        platform->setPositions(*this, p);
    }
    void getPositions(std::vector<Vec3> &out) const { platform->getPositions(*this, out); }
    const std::vector<Vec3> &positionsRef() const { return positions; }
    bool hasPositions() const { return positionsSet; }
    void getPeriodicBoxVectors(Vec3 &a, Vec3 &b, Vec3 &c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setPeriodicBoxVectors(const Vec3 &a, const Vec3 &b, const Vec3 &c) {
        if (a[0] != box[0][0] || a[1] != box[0][1] || a[2] != box[0][2] || b[0] != box[1][0] || b[1] != box[1][1] ||
            b[2] != box[1][2] || c[0] != box[2][0] || c[1] != box[2][1] || c[2] != box[2][2])
            boxVersion++;
        box[0] = a; box[1] = b; box[2] = c;
    }
    unsigned long getBoxVersion() const { return boxVersion; }
    /** Energy (kJ/mol) of the force groups in the bit mask `groups`; forces are left where the platform keeps them. */
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = -1);
    /** Host force accumulator of the current evaluation (kJ/mol/nm, HostB200 platform). */
    std::vector<Vec3> &getForces() { return forces; }
    /** A new Context on the same platform that shares this one's device and stream (OpenMM 7.7+). */
    Context *createLinkedContext(const System &system, Integrator &integrator);
    ForceImpl &getForceImpl(const Force &force) const {
        for (auto &fi : forceImpls)
            if (fi.first == &force) return *fi.second;
        throw OpenMMException("The Force is not part of this Context's System");
    }

private:
    friend class Context;
    Context &owner;
    const System &system;
    Platform *platform;
    void *platformData;
    std::map<std::string, double> parameters;
    std::vector<Vec3> positions, forces;
    Vec3 box[3];
    double time;
    bool positionsSet;
    unsigned long boxVersion;
    std::vector<std::pair<const Force *, std::unique_ptr<ForceImpl>>> forceImpls;
};

/** Owns the ContextImpl; evaluates force groups the way Context::getState does. */
class Context {
public:
    explicit Context(const System &system) : Context(system, HostPlatform::instance()) {}
    Context(const System &system, Platform &platform, const std::map<std::string, std::string> &properties = {})
        : impl(new ContextImpl(*this, system, platform, properties, nullptr)) {}
    const System &getSystem() const { return impl->getSystem(); }
    Platform &getPlatform() { return impl->getPlatform(); }
    void setPositions(const std::vector<Vec3> &positions) { impl->setPositions(positions); }
    void setPeriodicBoxVectors(const Vec3 &a, const Vec3 &b, const Vec3 &c) { impl->setPeriodicBoxVectors(a, b, c); }
    void setParameter(const std::string &name, double value) { impl->setParameter(name, value); }
    double getParameter(const std::string &name) const { return impl->getParameter(name); }
    const std::map<std::string, double> &getParameters() const { return impl->getParameters(); }
    void setTime(double t) { impl->setTime(t); }
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = -1) {
        if (!impl->hasPositions()) throw OpenMMException("Particle positions have not been set");
        const double e = impl->calcForcesAndEnergy(includeForces, includeEnergy, groups);
        impl->getPlatform().getForces(*impl, lastForces);
        return e;
    }
    const std::vector<Vec3> &getForces() const { return lastForces; }
    ContextImpl &getImpl() { return *impl; }
    const ContextImpl &getImpl() const { return *impl; }
    ForceImpl &getForceImpl(const Force &force) const { return impl->getForceImpl(force); }

private:
    friend class ContextImpl;
    Context(const System &system, Platform &platform, ContextImpl &original)
        : impl(new ContextImpl(*this, system, platform, {}, &original)) {}
    std::unique_ptr<ContextImpl> impl;
    std::vector<Vec3> lastForces;
};

inline ContextImpl &getContextImpl(Context &context) { return context.getImpl(); }   // OpenMM: ForceImpl::getContextImpl

}  // namespace OpenMM

#endif
