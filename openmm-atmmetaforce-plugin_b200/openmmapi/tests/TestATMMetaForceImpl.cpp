// Host logic of ATMMetaForceImpl and of the three Context-taking members of ATMMetaForce, in C++ against the
// OpenMM-free System / Context of this build.  The reference has no C++ test of its Impl (the
// per-platform tests are commented out); this one pins the behaviour its callers rely on
// (ref: openmmapi/src/ATMMetaForceImpl.cpp:69-88 mask rule, :103 group skip, :130-142 parameter names).
// Runs without a GPU; with one (argv[1] == "gpu") it also evaluates a small periodic system twice and checks that the
// force is translation invariant and that a zero displacement gives u = 0.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "ATMMetaForce.h"
#include "ATMMetaForceImpl.h"

using namespace ATMMetaForcePlugin;
using OpenMM::OpenMMException;

#define EXPECT(cond)                                                                \
    do {                                                                            \
        if (!(cond)) {                                                              \
            std::printf("FAILED %s (%s:%d)\n", #cond, __FILE__, __LINE__);         \
            return 1;                                                               \
        }                                                                           \
    } while (0)

template <class F>
static bool throwsWith(F f, const char *needle) {
    try {
        f();
    } catch (const OpenMMException &e) {
        return std::strstr(e.what(), needle) != nullptr;
    }
    return false;
}

// n^3 "ions" on a jittered cubic lattice, alternating charges, a 4-atom "ligand" displaced by half the box
static ATMMetaForce *fill(OpenMM::System &system, int n, double L, int atmGroup, std::vector<int> varGroups, int nbGroup,
                          std::vector<OpenMM::Vec3> &pos) {
    auto *nb = new OpenMM::NonbondedForce();
    unsigned s = 12345u;
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
    nb->setForceGroup(nbGroup);
    system.addForce(nb);
    system.setDefaultPeriodicBoxVectors(OpenMM::Vec3(L, 0, 0), OpenMM::Vec3(0, L, 0), OpenMM::Vec3(0, 0, L));
    auto *force = new ATMMetaForce(0.2, 0.7, 0.1, 5.0, 0.5, 800.0, 400.0, 0.0625, 1.0, varGroups);
    force->setForceGroup(atmGroup);
    for (int i = 0; i < n * n * n; i++) force->addParticle(i, i < 4 ? 0.5 * L : 0.0, 0.0, 0.0);
    system.addForce(force);
    return force;
}

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    {   // parameters, names, errors
        OpenMM::System system;
        std::vector<OpenMM::Vec3> pos;
        ATMMetaForce *force = fill(system, 12, 4.8, 3, {1}, 1, pos);
        OpenMM::Context context(system);
        EXPECT(context.getParameters().size() == 9);
        EXPECT(context.getParameter(ATMMetaForce::Lambda2()) == 0.7 && context.getParameter(ATMMetaForce::Umax()) == 800.0);
        context.setParameter(ATMMetaForce::Alpha(), 0.0);
        EXPECT(context.getParameter("ATMAlpha") == 0.0);
        EXPECT(throwsWith([&] { context.setParameter("ATMGamma", 1.0); }, "invalid parameter name"));
        EXPECT(throwsWith([&] { context.calcForcesAndEnergy(true, true); }, "positions have not been set"));
        EXPECT(force->getPerturbationEnergy(context) == 0.0);
        auto &impl = dynamic_cast<ATMMetaForceImpl &>(context.getForceImpl(*force));
        EXPECT(impl.getKernelNames().size() == 1 && impl.getKernelNames()[0] == "CalcATMMetaForce");
        EXPECT(impl.getVariableForceGroupsMask() == 2);
        EXPECT(&impl.getOwner() == force);
        context.setPositions(pos);
        EXPECT(context.calcForcesAndEnergy(true, true, 1 << 1) == 0.0);   // ATM group not requested: no work at all
        force->updateParametersInContext(context);
        if (!gpu) {
            EXPECT(throwsWith([&] { context.calcForcesAndEnergy(true, true); }, "CUDA"));   // no CPU fallback
        } else {
            const double e1 = context.calcForcesAndEnergy(true, true);
            const double u1 = force->getPerturbationEnergy(context);
            std::vector<OpenMM::Vec3> f1 = context.getForces();
            EXPECT(std::isfinite(e1) && std::isfinite(u1) && u1 != 0.0);
            // rigid translation by a lattice-incommensurate vector: same energy and forces (periodic system)
            std::vector<OpenMM::Vec3> moved = pos;
            for (auto &p : moved) { p[0] += 0.137; p[1] -= 0.291; p[2] += 1.013; }
            context.setPositions(moved);
            const double e2 = context.calcForcesAndEnergy(true, true);
            EXPECT(std::fabs(e2 - e1) <= 2e-5 * std::fabs(e1) + 1e-2);
            double num = 0, den = 0;
            for (size_t i = 0; i < f1.size(); i++)
                for (int c = 0; c < 3; c++) {
                    const double d = context.getForces()[i][c] - f1[i][c];
                    num += d * d; den += f1[i][c] * f1[i][c];
                }
            EXPECT(std::sqrt(num / den) <= 1e-4);   // float coordinates: the translation itself rounds at ~1e-7 nm
            // zero displacement: both states coincide, u == 0 exactly
            for (int i = 0; i < 4; i++) force->setParticleParameters(i, i, 0.0, 0.0, 0.0);
            force->updateParametersInContext(context);
            context.calcForcesAndEnergy(true, true);
            EXPECT(force->getPerturbationEnergy(context) == 0.0);
        }
    }
    {   // the ATM group cannot be variable; particle counts must agree
        OpenMM::System system;
        std::vector<OpenMM::Vec3> pos;
        fill(system, 4, 4.0, 1, {1}, 2, pos);
        EXPECT(throwsWith([&] { OpenMM::Context c(system); }, "cannot be one of the variable force groups"));
        OpenMM::System system2;
        std::vector<OpenMM::Vec3> pos2;
        fill(system2, 4, 4.0, 3, {1}, 1, pos2);
        system2.addParticle(1.0);
        EXPECT(throwsWith([&] { OpenMM::Context c(system2); }, "exactly as many particles"));
    }
    std::printf("Done\n");
    return 0;
}
