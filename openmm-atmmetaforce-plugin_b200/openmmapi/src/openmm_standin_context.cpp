// openmm_standin_context.cpp -- the non-inline part of the OpenMM-free Context / Platform stand-ins (see
// openmm_standin_context.h).  Not compiled with -DATM_HAVE_OPENMM.
#ifndef ATM_HAVE_OPENMM
#include "openmm_standin_context.h"

#include <dlfcn.h>

namespace OpenMM {

// ---- Platform: default hooks = host vectors ---------------------------------------------------------------------
void Platform::contextCreated(ContextImpl &, const std::map<std::string, std::string> &) const {}
void Platform::linkedContextCreated(ContextImpl &, ContextImpl &) const {}
void Platform::contextDestroyed(ContextImpl &) const {}
void Platform::setPositions(ContextImpl &, const std::vector<Vec3> &) const {}
void Platform::getPositions(const ContextImpl &context, std::vector<Vec3> &positions) const { positions = context.positionsRef(); }
void Platform::beginComputation(ContextImpl &context, bool, bool, int) const {
    context.getForces().assign(context.getSystem().getNumParticles(), Vec3());
}
double Platform::finishComputation(ContextImpl &, bool, bool, int) const { return 0.0; }
void Platform::getForces(ContextImpl &context, std::vector<Vec3> &forces) const { forces = context.getForces(); }
void Platform::addForces(ContextImpl &context, const std::vector<Vec3> &forces) const {
    std::vector<Vec3> &f = context.getForces();
    for (size_t i = 0; i < forces.size(); i++)
        for (int c = 0; c < 3; c++) f[i][c] += forces[i][c];
}

std::vector<std::unique_ptr<Platform>> &Platform::registry() {
    static std::vector<std::unique_ptr<Platform>> r;
    return r;
}

HostPlatform &HostPlatform::instance() {
    static HostPlatform p;
    return p;
}

void Platform::loadPluginLibrary(const std::string &file) {
    void *lib = dlopen(file.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw OpenMMException("Error loading library " + file + ": " + dlerror());
    for (const char *sym : {"registerPlatforms", "registerKernelFactories"}) {
        void (*fn)() = reinterpret_cast<void (*)()>(dlsym(lib, sym));
        if (fn) fn();
    }
}

// ---- ContextImpl ------------------------------------------------------------------------------------------------
ContextImpl::ContextImpl(Context &owner, const System &system, Platform &platform, const std::map<std::string, std::string> &properties,
                         ContextImpl *originalContext)
    : owner(owner), system(system), platform(&platform), platformData(nullptr), time(0.0), positionsSet(false), boxVersion(0) {
    system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
    if (originalContext) platform.linkedContextCreated(*this, *originalContext);
    else platform.contextCreated(*this, properties);
    try {
        for (int i = 0; i < system.getNumForces(); i++) {
            ForceImpl *fi = system.getForce(i).createImpl();
            if (!fi) continue;
            forceImpls.emplace_back(&system.getForce(i), std::unique_ptr<ForceImpl>(fi));
            for (const auto &kv : fi->getDefaultParameters()) parameters[kv.first] = kv.second;
        }
        for (auto &fi : forceImpls) fi.second->initialize(*this);
    } catch (...) {
        forceImpls.clear();
        platform.contextDestroyed(*this);
        throw;
    }
}

ContextImpl::~ContextImpl() {
    forceImpls.clear();   // kernels hold references into the platform data
    platform->contextDestroyed(*this);
}

double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups) {
    platform->beginComputation(*this, includeForces, includeEnergy, groups);
    double energy = 0.0;
    for (auto &fi : forceImpls) energy += fi.second->calcForcesAndEnergy(*this, includeForces, includeEnergy, groups);
    energy += platform->finishComputation(*this, includeForces, includeEnergy, groups);
    return energy;
}

Context *ContextImpl::createLinkedContext(const System &innerSystem, Integrator &) {
    return new Context(innerSystem, *platform, *this);
}

// ---- NonbondedForce as an ordinary Force with an Impl (platforms that have a CalcNonbondedForce kernel) ---------------
namespace {
class NonbondedForceImpl : public ForceImpl {
public:
    explicit NonbondedForceImpl(const NonbondedForce &owner) : owner(owner) {}
    void initialize(ContextImpl &context) override {
        if (!context.getPlatform().supportsKernels({CalcNonbondedForceKernel::Name()})) return;   // host platform: data only
        if (owner.getNumParticles() != context.getSystem().getNumParticles())
            throw OpenMMException("NonbondedForce must have exactly as many particles as the System it belongs to.");
        kernel = context.getPlatform().createKernel(CalcNonbondedForceKernel::Name(), context);
        kernel.getAs<CalcNonbondedForceKernel>().initialize(context.getSystem(), owner);
    }
    double calcForcesAndEnergy(ContextImpl &context, bool includeForces, bool includeEnergy, int groups) override {
        if (!kernel) return 0.0;
        const bool direct = (groups & (1 << owner.getForceGroup())) != 0;
        const int rg = owner.getReciprocalSpaceForceGroup();
        const bool recip = rg < 0 ? direct : (groups & (1 << rg)) != 0;
        if (!direct && !recip) return 0.0;
        return kernel.getAs<CalcNonbondedForceKernel>().execute(context, includeForces, includeEnergy, direct, recip);
    }
    std::map<std::string, double> getDefaultParameters() override { return {}; }
    std::vector<std::string> getKernelNames() override { return {CalcNonbondedForceKernel::Name()}; }

private:
    const NonbondedForce &owner;
    Kernel kernel;
};
}  // namespace

ForceImpl *NonbondedForce::createImpl() const { return new NonbondedForceImpl(*this); }

// ---- host-evaluated forces ---------------------------------------------------------------------------------------
namespace {
class HostEvaluatedForceImpl : public ForceImpl {
public:
    explicit HostEvaluatedForceImpl(const HostEvaluatedForce &owner) : owner(owner) {}
    void initialize(ContextImpl &) override {}
    double calcForcesAndEnergy(ContextImpl &context, bool includeForces, bool includeEnergy, int groups) override {
        if ((groups & (1 << owner.getForceGroup())) == 0) return 0.0;
        std::vector<Vec3> pos;
        context.getPositions(pos);
        Vec3 box[3];
        context.getPeriodicBoxVectors(box[0], box[1], box[2]);
        std::vector<Vec3> f;
        if (includeForces) f.assign(pos.size(), Vec3());
        const double e = owner.evaluate(pos, box, includeForces ? &f : nullptr);
        if (includeForces) context.getPlatform().addForces(context, f);
        return includeEnergy ? e : 0.0;
    }
    std::map<std::string, double> getDefaultParameters() override { return {}; }
    std::vector<std::string> getKernelNames() override { return {}; }

private:
    const HostEvaluatedForce &owner;
};
}  // namespace

ForceImpl *HostEvaluatedForce::createImpl() const { return new HostEvaluatedForceImpl(*this); }

double HarmonicBondForce::evaluate(const std::vector<Vec3> &positions, const Vec3 box[3], std::vector<Vec3> *forces) const {
    double energy = 0.0;
    for (const Bond &b : bonds) {
        if (b.p1 < 0 || b.p2 < 0 || b.p1 >= (int)positions.size() || b.p2 >= (int)positions.size())
            throw OpenMMException("HarmonicBondForce: Illegal particle index for a bond");
        double d[3];
        for (int c = 0; c < 3; c++) d[c] = positions[b.p2][c] - positions[b.p1][c];
        if (periodic)   // rectangular boxes only, like the rest of this build
            for (int c = 0; c < 3; c++) d[c] -= box[c][c] * std::floor(d[c] / box[c][c] + 0.5);
        const double r = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double dr = r - b.length;
        energy += 0.5 * b.k * dr * dr;
        if (forces && r > 0.0) {
            const double s = b.k * dr / r;   // dE/dr / r
            for (int c = 0; c < 3; c++) {
                (*forces)[b.p1][c] += s * d[c];
                (*forces)[b.p2][c] -= s * d[c];
            }
        }
    }
    return energy;
}

void pmeGridDimensions(double alpha, double tolerance, const Vec3 box[3], int grid[3]) {
    for (int k = 0; k < 3; k++) {
        int n = (int)std::ceil(2.0 * alpha * box[k][k] / (3.0 * std::pow(tolerance, 0.2)));
        n = std::max(n, 6);
        for (;; n++) {   // next size whose only prime factors are 2, 3, 5, 7
            int m = n;
            for (int f : {2, 3, 5, 7})
                while (m % f == 0) m /= f;
            if (m == 1) break;
        }
        grid[k] = n;
    }
}

}  // namespace OpenMM
#endif
