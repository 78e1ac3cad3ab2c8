// B200ATMMetaForceKernels.h -- the Blackwell implementation of the "CalcATMMetaForce" kernel for OpenMM's CUDA platform.
// It replaces platforms/common + platforms/cuda (+ hip, opencl) of the reference: where those JIT-compile the two device
// kernels of kernels/atmmetaforce.cc and bind OpenMM's arrays as arguments
// (ref: platforms/common/src/CommonATMMetaForceKernels.cpp:111-152, platforms/cuda/include/CudaATMMetaForceKernels.h:13-20),
// this class hands the same borrowed device pointers -- posq, posqCorrection, the long force buffers of the outer and the
// two inner contexts, on the context's own stream -- to the pre-compiled sm_100a library through ATMMetaForceB200Kernel.
// Compiles against OpenMM (-DATM_HAVE_OPENMM) or against the stand-ins of this repository.
#ifndef B200_ATMMETAFORCE_KERNELS_H_
#define B200_ATMMETAFORCE_KERNELS_H_

#include "ATMMetaForceB200Kernel.h"
#include "ATMMetaForceKernels.h"
#ifdef ATM_HAVE_OPENMM
#include "openmm/common/ContextSelector.h"
#include "openmm/cuda/CudaContext.h"
#include "openmm/cuda/CudaPlatform.h"
#else
#include "openmm_standin_cuda.h"
#endif

namespace ATMMetaForcePlugin {

class B200CalcATMMetaForceKernel : public CalcATMMetaForceKernel {
public:
    B200CalcATMMetaForceKernel(std::string name, const OpenMM::Platform &platform, OpenMM::CudaContext &cu)
        : CalcATMMetaForceKernel(name, platform), cu(cu), owner(nullptr), hasListener(false) {}
    /** ref: CommonCalcATMMetaForceKernel::initialize (CommonATMMetaForceKernels.cpp:78-109). */
    void initialize(const OpenMM::System &system, const ATMMetaForce &force) override;
    /** ref: CommonCalcATMMetaForceKernel::execute (:154-204). */
    double execute(OpenMM::ContextImpl &context, OpenMM::ContextImpl &innerContext1, OpenMM::ContextImpl &innerContext2, double State1Energy,
                   double State2Energy, bool includeForces, bool includeEnergy) override;
    /** ref: CommonCalcATMMetaForceKernel::copyState (:206-226). */
    void copyState(OpenMM::ContextImpl &context, OpenMM::ContextImpl &innerContext1, OpenMM::ContextImpl &innerContext2) override;
    /** ref: CommonCalcATMMetaForceKernel::copyParametersToContext (:229-251). */
    void copyParametersToContext(OpenMM::ContextImpl &context, const ATMMetaForce &force) override;
    double getPerturbationEnergy() override { return kernel.getPerturbationEnergy(); }
    /** ref: CudaCalcATMMetaForceKernel::getInnerComputeContext (platforms/cuda/include/CudaATMMetaForceKernels.h:17-19). */
    OpenMM::CudaContext &getInnerComputeContext(OpenMM::ContextImpl &innerContext);
    ATMMetaForceB200Kernel &getBackend() { return kernel; }

private:
    class ReorderListener;
    OpenMM::CudaContext &cu;
    const ATMMetaForce *owner;
    bool hasListener;
    ATMMetaForceB200Kernel kernel;
};

}  // namespace ATMMetaForcePlugin

#endif
