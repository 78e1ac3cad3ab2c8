// XML round trip of ATMMetaForce with the values of the reference's own serialization test
// (ref: serialization/tests/TestSerializeATMMetaForce.cpp:13-24,40-41: name MyATMMetaForce, group 30,
// (0, .1, .25, .5, .6, 200, 100, .07, 0), groups {1}, particles (0; .1,.2,.3) and (2; .4,.5,.6)).
#include <cstdio>
#include <iostream>
#include <sstream>

#include "ATMMetaForce.h"
#include "ATMMetaForceB200Kernel.h"
#include "ATMMetaForceProxy.h"

using namespace ATMMetaForcePlugin;

#define EXPECT(cond)                                                                \
    do {                                                                            \
        if (!(cond)) {                                                              \
            std::printf("FAILED %s (%s:%d)\n", #cond, __FILE__, __LINE__);         \
            return 1;                                                               \
        }                                                                           \
    } while (0)

int main(int argc, char **argv) {
    ATMMetaForce force(0.0, 0.1, 0.25, 0.5, 0.6, 200.0, 100.0, 0.07, 0.0, std::vector<int>{1});
    force.setForceGroup(30);
    force.setName("MyATMMetaForce");
    force.addParticle(0, 0.1, 0.2, 0.3);
    force.addParticle(2, 0.4, 0.5, 0.6);

    std::stringstream buffer;
    OpenMM::XmlSerializer::serialize<ATMMetaForce>(&force, "Force", buffer, "ATMMetaForce");
    if (argc > 1) std::cout << buffer.str();
    ATMMetaForce *copy = OpenMM::XmlSerializer::deserialize<ATMMetaForce>(buffer);

    EXPECT(copy->getForceGroup() == 30);
    EXPECT(copy->getName() == "MyATMMetaForce");
    EXPECT(copy->getDefaultLambda1() == 0.0 && copy->getDefaultLambda2() == 0.1 && copy->getDefaultAlpha() == 0.25);
    EXPECT(copy->getDefaultU0() == 0.5 && copy->getDefaultW0() == 0.6 && copy->getDefaultUmax() == 200.0);
    EXPECT(copy->getDefaultUbcore() == 100.0 && copy->getDefaultAcore() == 0.07 && copy->getDefaultDirection() == 0.0);
    EXPECT(copy->getVariableForceGroups() == std::vector<int>{1});
    EXPECT(copy->getNumParticles() == 2);
    int particle;
    double dx, dy, dz;
    copy->getParticleParameters(0, particle, dx, dy, dz);
    EXPECT(particle == 0 && dx == 0.1 && dy == 0.2 && dz == 0.3);
    copy->getParticleParameters(1, particle, dx, dy, dz);
    EXPECT(particle == 2 && dx == 0.4 && dy == 0.5 && dz == 0.6);  // the particle field is stored verbatim

    // error behaviour
    bool threw = false;
    try { copy->getParticleParameters(2, particle, dx, dy, dz); } catch (const OpenMM::OpenMMException &) { threw = true; }
    EXPECT(threw);
    threw = false;
    try {
        std::stringstream bad("<Force type=\"ATMMetaForce\" version=\"1\"/>");
        OpenMM::XmlSerializer::deserialize<ATMMetaForce>(bad);
    } catch (const OpenMM::OpenMMException &e) { threw = std::string(e.what()) == "Unsupported version"; }
    EXPECT(threw);
    threw = false;
    try {
        ATMMetaForce f2(0, 0, 0, 0, 0, 1, 1, 1, 1, std::vector<int>{3});
        f2.setForceGroup(3);
        variableForceGroupsMask(f2);
    } catch (const OpenMM::OpenMMException &) { threw = true; }
    EXPECT(threw);
    EXPECT(variableForceGroupsMask(force) == 2);
    auto defaults = ATMMetaForceB200Kernel::getDefaultParameters(force);
    EXPECT(defaults.size() == 9 && defaults["ATMLambda2"] == 0.1 && defaults["ATMUmax"] == 200.0);
    delete copy;
    std::printf("Done\n");
    return 0;
}
