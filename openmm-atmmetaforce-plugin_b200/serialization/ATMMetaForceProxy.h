#ifndef OPENMM_ATMMETAFORCE_PROXY_H_
#define OPENMM_ATMMETAFORCE_PROXY_H_
#ifdef ATM_HAVE_OPENMM
#include "openmm/serialization/SerializationProxy.h"
#else
#include "openmm_standin.h"
#endif

namespace ATMMetaForcePlugin {

/** XML (de)serialisation of ATMMetaForce.  Schema identical to the reference
 *  (ref: serialization/src/ATMMetaForceProxy.cpp:11-40): element type="ATMMetaForce" version="0", attributes
 *  forceGroup name lambda1 lambda2 alpha u0 w0 uMax ubCore aCore direction, children
 *  <VariableForceGroups><Parameter group=.../></VariableForceGroups> and
 *  <Particles><Particle particle dx dy dz/></Particles>. */
class ATMMetaForceProxy : public OpenMM::SerializationProxy {
public:
    ATMMetaForceProxy();
    void serialize(const void *object, OpenMM::SerializationNode &node) const;
    void *deserialize(const OpenMM::SerializationNode &node) const;
};

/** Registers the proxy (the reference does this from a static constructor of libATMMetaForcePlugin.so,
 *  ref: serialization/src/SerializationProxyRegistration.cpp:15-22). */
void registerATMMetaForceSerializationProxies();

}  // namespace ATMMetaForcePlugin
#endif
