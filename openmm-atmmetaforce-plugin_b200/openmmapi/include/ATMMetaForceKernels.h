// ATMMetaForceKernels.h -- the kernel seam between ATMMetaForceImpl and a platform back-end: an OpenMM::KernelImpl named
// "CalcATMMetaForce" with five operations.  Same name, methods and argument meaning as the reference's seam
// (ref: openmmapi/include/ATMMetaForceKernels.h:15-59), so that a platform library written against one header works
// with the other.  The Blackwell implementation is platforms/b200/include/B200ATMMetaForceKernels.h.
#ifndef ATMMETAFORCE_KERNELS_H_
#define ATMMETAFORCE_KERNELS_H_

#include <string>

#include "ATMMetaForce.h"
#ifdef ATM_HAVE_OPENMM
#include "openmm/KernelImpl.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"
#else
#include "openmm_standin_context.h"
#endif

namespace ATMMetaForcePlugin {

class CalcATMMetaForceKernel : public OpenMM::KernelImpl {
public:
    static std::string Name() { return "CalcATMMetaForce"; }
    CalcATMMetaForceKernel(std::string name, const OpenMM::Platform &platform) : OpenMM::KernelImpl(name, platform) {}
    /** Called once from ATMMetaForceImpl::initialize: displacement table, reorder listener. */
    virtual void initialize(const OpenMM::System &system, const ATMMetaForce &force) = 0;
    /** Scalar stage (u, soft core, softplus, sp) from the two inner energies, then force += sp F2 + (1 - sp) F1.
     *  Returns e0 + W when includeEnergy, else 0. */
    virtual double execute(OpenMM::ContextImpl &context, OpenMM::ContextImpl &innerContext1, OpenMM::ContextImpl &innerContext2,
                           double State1Energy, double State2Energy, bool includeForces, bool includeEnergy) = 0;
    /** posq1 = posq, posq2 = posq + displacement; box vectors, time and global parameters mirrored into the inner contexts. */
    virtual void copyState(OpenMM::ContextImpl &context, OpenMM::ContextImpl &innerContext1, OpenMM::ContextImpl &innerContext2) = 0;
    /** Re-reads the displacement vectors of the force (ATMMetaForce::updateParametersInContext). */
    virtual void copyParametersToContext(OpenMM::ContextImpl &context, const ATMMetaForce &force) = 0;
    /** u_sc of the last execute(). */
    virtual double getPerturbationEnergy() = 0;
};

}  // namespace ATMMetaForcePlugin

#endif
