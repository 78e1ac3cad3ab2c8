#include "ATMMetaForce.h"

using namespace ATMMetaForcePlugin;

int ATMMetaForce::addParticle(int particle, double dx, double dy, double dz) {
    particles.push_back(ParticleInfo{particle, dx, dy, dz});
    return (int)particles.size() - 1;
}

void ATMMetaForce::getParticleParameters(int index, int &particle, double &dx, double &dy, double &dz) const {
    ASSERT_VALID_INDEX(index, particles);
    const ParticleInfo &p = particles[index];
    particle = p.particle;
    dx = p.dx;
    dy = p.dy;
    dz = p.dz;
}

void ATMMetaForce::setParticleParameters(int index, int particle, double dx, double dy, double dz) {
    ASSERT_VALID_INDEX(index, particles);
    particles[index] = ParticleInfo{particle, dx, dy, dz};
}

void ATMMetaForce::getDefaultParameters(double p[9]) const {
    p[0] = defaultLambda1; p[1] = defaultLambda2; p[2] = defaultAlpha; p[3] = defaultU0; p[4] = defaultW0;
    p[5] = defaultUmax; p[6] = defaultUbcore; p[7] = defaultAcore; p[8] = defaultDirection;
}

std::vector<double> ATMMetaForce::getDisplacementArray() const {
    std::vector<double> d(3 * particles.size());
    for (size_t i = 0; i < particles.size(); i++) {
        d[3 * i] = particles[i].dx;
        d[3 * i + 1] = particles[i].dy;
        d[3 * i + 2] = particles[i].dz;
    }
    return d;
}
