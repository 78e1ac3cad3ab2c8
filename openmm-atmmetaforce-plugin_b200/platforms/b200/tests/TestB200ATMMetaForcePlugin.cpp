// The compiled plugin glue, exercised the way OpenMM exercises a plugin: the library is dlopen()ed, its three
// extern "C" registration symbols are looked up (ref: platforms/cuda/src/CudaATMMetaForceKernelFactory.cpp:14-36),
// registerATMMetaForceCudaKernelFactories() registers the CUDA platform when it is absent and the kernel factory on it,
// and Platform::createKernel("CalcATMMetaForce") then serves ATMMetaForceImpl.  Without a GPU the test stops after the
// registration checks (creating a Context on the CUDA platform must then fail with the device error).  With one
// (argv[1] == "gpu") it runs the reference's orchestration end to end on a small periodic system -- copyState into two
// linked inner contexts, the two inner evaluations of the variable force group, execute -- and compares energy,
// perturbation energy and forces with the fused Tier-2 path of the kernel-less host platform on the same System.
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "ATMMetaForce.h"
#include "ATMMetaForceImpl.h"
#include "ATMMetaForceKernels.h"
#include "openmm_standin_cuda.h"

using namespace ATMMetaForcePlugin;
using OpenMM::OpenMMException;

#define EXPECT(cond)                                                                \
    do {                                                                            \
        if (!(cond)) {                                                              \
            std::printf("FAILED %s (%s:%d)\n", #cond, __FILE__, __LINE__);         \
            return 1;                                                               \
        }                                                                           \
    } while (0)

static ATMMetaForce *fill(OpenMM::System &system, int n, double L, std::vector<OpenMM::Vec3> &pos) {
    auto *nb = new OpenMM::NonbondedForce();
    unsigned s = 777u;
    auto rnd = [&s]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / 16777216.0 - 0.5; };
    const double a = L / n;
    for (int i = 0; i < n * n * n; i++) {
        system.addParticle(16.0);
        nb->addParticle((i % 2) ? 0.4 : -0.4, 0.3, 0.6);
        pos.push_back(OpenMM::Vec3((i % n + 0.5 + 0.3 * rnd()) * a, ((i / n) % n + 0.5 + 0.3 * rnd()) * a, (i / (n * n) + 0.5 + 0.3 * rnd()) * a));
    }
    nb->addException(0, 1, 0.0, 0.3, 0.0);
    nb->addException(2, 3, -0.05, 0.3, 0.2);
    nb->setCutoffDistance(0.9);
    nb->setForceGroup(1);
    system.addForce(nb);
    system.setDefaultPeriodicBoxVectors(OpenMM::Vec3(L, 0, 0), OpenMM::Vec3(0, L, 0), OpenMM::Vec3(0, 0, L));
    auto *force = new ATMMetaForce(0.2, 0.7, 0.1, 5.0, 0.5, 800.0, 400.0, 0.0625, 1.0, {1});
    force->setForceGroup(3);
    for (int i = 0; i < n * n * n; i++) force->addParticle(i, i < 4 ? 0.5 * L : 0.0, 0.0, 0.0);
    system.addForce(force);
    return force;
}

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    const std::string dir = argc > 2 ? argv[2] : ".";
    const std::string lib = dir + "/libATMMetaForcePluginCUDA.so";

    // ---- the plugin library and its symbols
    void *h = dlopen(lib.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!h) std::printf("dlopen: %s\n", dlerror());
    EXPECT(h != nullptr);
    void (*regPlatforms)() = reinterpret_cast<void (*)()>(dlsym(h, "registerPlatforms"));
    void (*regFactories)() = reinterpret_cast<void (*)()>(dlsym(h, "registerKernelFactories"));
    void (*regCuda)() = reinterpret_cast<void (*)()>(dlsym(h, "registerATMMetaForceCudaKernelFactories"));
    EXPECT(regPlatforms && regFactories && regCuda);

    // registerKernelFactories without a CUDA platform: silently nothing (ref: the catch block at :23-25)
    const int before = OpenMM::Platform::getNumPlatforms();
    regPlatforms();
    regFactories();
    EXPECT(OpenMM::Platform::getNumPlatforms() == before);
    bool threw = false;
    try { OpenMM::Platform::getPlatformByName("CUDA"); } catch (const OpenMMException &) { threw = true; }
    EXPECT(threw);
    // the static-link entry point registers the platform when absent (ref: :28-36), and is idempotent about that
    regCuda();
    EXPECT(OpenMM::Platform::getNumPlatforms() == before + 1);
    OpenMM::Platform &cuda = OpenMM::Platform::getPlatformByName("CUDA");
    EXPECT(cuda.supportsKernels({CalcATMMetaForceKernel::Name()}));
    EXPECT(!cuda.supportsKernels({"CalcSomethingElse"}));
    regCuda();
    EXPECT(OpenMM::Platform::getNumPlatforms() == before + 1);
    EXPECT(!OpenMM::HostPlatform::instance().supportsKernels({CalcATMMetaForceKernel::Name()}));

    OpenMM::System system;
    std::vector<OpenMM::Vec3> pos;
    ATMMetaForce *force = fill(system, 12, 4.8, pos);
    if (!gpu) {
        threw = false;
        try {
            OpenMM::Context c(system, cuda);
        } catch (const OpenMMException &e) {
            threw = std::strstr(e.what(), "CUDA") != nullptr;
        }
        EXPECT(threw);   // no device: no Context, no CPU fallback
        std::printf("Done\n");
        return 0;
    }

    // ---- the reference's orchestration on the CUDA platform
    OpenMM::Context context(system, cuda, {{"Precision", "mixed"}});
    auto &impl = dynamic_cast<ATMMetaForceImpl &>(context.getForceImpl(*force));
    EXPECT(impl.usesPlatformKernel());
    context.setPositions(pos);
    EXPECT(context.calcForcesAndEnergy(true, true, 1 << 0) == 0.0);       // neither the ATM group nor the variable group
    const double eATM = context.calcForcesAndEnergy(true, true, 1 << 3);   // the ATM group alone
    const double uATM = force->getPerturbationEnergy(context);
    std::vector<OpenMM::Vec3> fATM = context.getForces();
    EXPECT(std::isfinite(eATM) && std::isfinite(uATM) && uATM != 0.0);
    EXPECT(impl.getInnerContext(1) && impl.getInnerContext(2));
    EXPECT(impl.getInnerContext(1)->getSystem().getNumForces() == 1);      // the NonbondedForce clone; the ATM force stays out

    // ---- the same System through the fused Tier-2 path (kernel-less host platform, direct space only): the inner
    //      contexts of the CUDA stand-in evaluate direct + reciprocal space + dispersion correction, so U1 / U2 differ
    //      by those terms while u = U2 - U1 differs by the reciprocal-space DIFFERENCE only; compare what must agree --
    //      the soft-core / softplus relation between the recorded energies, and the merge identity on the forces
    const std::vector<double> &rec = impl.getEnergyRecord();
    double out[7];
    const double p[ATM_NUM_PARAMS] = {0.2, 0.7, 0.1, 5.0, 0.5, 800.0, 400.0, 0.0625, 1.0};
    EXPECT(atm_softcore_softplus(p, rec[ATM_E_U1], rec[ATM_E_U2], out) == ATM_OK);
    // out = {u_sc, fp, ebias, bfp, energy, sp, sp_ref}
    EXPECT(std::fabs(out[0] - uATM) <= 1e-9 * std::fabs(uATM) + 1e-9);     // u_sc
    EXPECT(std::fabs(out[4] - eATM) <= 1e-9 * std::fabs(eATM));            // e0 + W
    // forces of the two inner contexts, blended on the host with sp, equal the outer force (2^-32 fixed-point rounding)
    const double sp = out[5];
    context.getImpl();   // (the inner force buffers are still valid: nothing ran since)
    std::vector<OpenMM::Vec3> f1, f2;
    OpenMM::ContextImpl &in1 = impl.getInnerContext(1)->getImpl(), &in2 = impl.getInnerContext(2)->getImpl();
    cuda.getForces(in1, f1);
    cuda.getForces(in2, f2);
    double worst = 0.0;
    for (size_t i = 0; i < fATM.size(); i++)
        for (int c = 0; c < 3; c++) worst = std::max(worst, std::fabs(fATM[i][c] - (sp * f2[i][c] + (1.0 - sp) * f1[i][c])));
    EXPECT(worst <= 2.0 / 4294967296.0 + 1e-12);

    // ---- reorder listener: a new atom order re-uploads the displacement table; results are unchanged
    std::vector<int> order(pos.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)((i * 7919u + 13u) % order.size());   // 7919 is coprime with 1728
    OpenMM::CudaPlatform::cudaContext(context.getImpl()).reorderAtoms(order);
    const double eATM2 = context.calcForcesAndEnergy(true, true, 1 << 3);
    EXPECT(std::fabs(eATM2 - eATM) <= 1e-9 * std::fabs(eATM));
    worst = 0.0;
    for (size_t i = 0; i < fATM.size(); i++)
        for (int c = 0; c < 3; c++) worst = std::max(worst, std::fabs(context.getForces()[i][c] - fATM[i][c]));
    EXPECT(worst <= 1e-6);

    // ---- updateParametersInContext through the seam: zero displacement -> both states coincide
    for (int i = 0; i < 4; i++) force->setParticleParameters(i, i, 0.0, 0.0, 0.0);
    force->updateParametersInContext(context);
    context.calcForcesAndEnergy(true, true, 1 << 3);
    EXPECT(force->getPerturbationEnergy(context) == 0.0);
    std::printf("Done\n");
    return 0;
}
