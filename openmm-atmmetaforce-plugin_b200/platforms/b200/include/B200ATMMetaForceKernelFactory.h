// B200ATMMetaForceKernelFactory.h -- creates the Blackwell "CalcATMMetaForce" kernel for contexts of the CUDA platform
// (ref: platforms/cuda/include/CudaATMMetaForceKernelFactory.h).
#ifndef B200_ATMMETAFORCE_KERNEL_FACTORY_H_
#define B200_ATMMETAFORCE_KERNEL_FACTORY_H_

#ifdef ATM_HAVE_OPENMM
#include "openmm/KernelFactory.h"
#else
#include "openmm_standin_context.h"
#endif

namespace ATMMetaForcePlugin {

class B200ATMMetaForceKernelFactory : public OpenMM::KernelFactory {
public:
    OpenMM::KernelImpl *createKernelImpl(std::string name, const OpenMM::Platform &platform, OpenMM::ContextImpl &context) const;
};

}  // namespace ATMMetaForcePlugin

// the three symbols OpenMM's plugin loader and the Python wrapper look for
// (ref: platforms/cuda/src/CudaATMMetaForceKernelFactory.cpp:14-36)
extern "C" {
void registerPlatforms();
void registerKernelFactories();
void registerATMMetaForceCudaKernelFactories();
}

#endif
