// B200ATMMetaForceKernels.cpp -- see B200ATMMetaForceKernels.h.
#include "B200ATMMetaForceKernels.h"

#include <map>

using namespace ATMMetaForcePlugin;
using namespace OpenMM;

// Re-uploads the displacement table in the new atom order whenever OpenMM re-sorts the atoms
// (ref: ReorderListener, CommonATMMetaForceKernels.cpp:46-72, hooked up at :120-122).
class B200CalcATMMetaForceKernel::ReorderListener : public CudaContext::ReorderListener {
public:
    explicit ReorderListener(B200CalcATMMetaForceKernel &k) : k(k) {}
    void execute() override { k.kernel.atomsReordered(*k.owner, k.cu.getAtomIndex(), k.cu.getCurrentStream()); }

private:
    B200CalcATMMetaForceKernel &k;
};

CudaContext &B200CalcATMMetaForceKernel::getInnerComputeContext(ContextImpl &innerContext) {
    return *static_cast<CudaPlatform::PlatformData *>(innerContext.getPlatformData())->contexts[0];
}

void B200CalcATMMetaForceKernel::initialize(const System &system, const ATMMetaForce &force) {
    ContextSelector selector(cu);
    owner = &force;
    if (force.getNumParticles() == 0) return;   // as the reference: nothing to set up (:81-82)
    if (force.getNumParticles() != system.getNumParticles())
        throw OpenMMException("ATMMetaForce must have exactly as many particles as the System it belongs to.");
    const atm_precision precision = cu.getUseDoublePrecision() ? ATM_PREC_DOUBLE : (cu.getUseMixedPrecision() ? ATM_PREC_MIXED : ATM_PREC_SINGLE);
    kernel.initialize(force, cu.getPaddedNumAtoms(), precision, cu.getAtomIndex(), cu.getDeviceIndex());
    if (!hasListener) {
        cu.addReorderListener(new ReorderListener(*this));   // owned by the context
        hasListener = true;
    }
}

void B200CalcATMMetaForceKernel::copyState(ContextImpl &context, ContextImpl &innerContext1, ContextImpl &innerContext2) {
    ContextSelector selector(cu);
    CudaContext &cu1 = getInnerComputeContext(innerContext1), &cu2 = getInnerComputeContext(innerContext2);
    if (owner && owner->getNumParticles() > 0) {
        const bool corr = cu.getUseMixedPrecision();
        kernel.copyState(cu.getPosq().getDevicePointer(), corr ? cu.getPosqCorrection().getDevicePointer() : nullptr,
                         cu1.getPosq().getDevicePointer(), corr ? cu1.getPosqCorrection().getDevicePointer() : nullptr,
                         cu2.getPosq().getDevicePointer(), corr ? cu2.getPosqCorrection().getDevicePointer() : nullptr, cu.getCurrentStream());
    }
    // host side of the state: box vectors, time and every global parameter the inner contexts know (:214-225)
    Vec3 a, b, c;
    context.getPeriodicBoxVectors(a, b, c);
    for (ContextImpl *inner : {&innerContext1, &innerContext2}) {
        inner->setPeriodicBoxVectors(a, b, c);
        inner->setTime(context.getTime());
        const std::map<std::string, double> innerParameters = inner->getParameters();
        for (const auto &param : innerParameters) inner->setParameter(param.first, context.getParameter(param.first));
    }
}

double B200CalcATMMetaForceKernel::execute(ContextImpl &context, ContextImpl &innerContext1, ContextImpl &innerContext2, double State1Energy,
                                           double State2Energy, bool includeForces, bool includeEnergy) {
    ContextSelector selector(cu);
    if (!owner || owner->getNumParticles() == 0) return 0.0;
    CudaContext &cu1 = getInnerComputeContext(innerContext1), &cu2 = getInnerComputeContext(innerContext2);
    std::map<std::string, double> parameters;
    for (const std::string *name : {&ATMMetaForce::Lambda1(), &ATMMetaForce::Lambda2(), &ATMMetaForce::Alpha(), &ATMMetaForce::U0(),
                                    &ATMMetaForce::W0(), &ATMMetaForce::Umax(), &ATMMetaForce::Ubcore(), &ATMMetaForce::Acore(),
                                    &ATMMetaForce::Direction()})
        parameters[*name] = context.getParameter(*name);
    return kernel.execute(parameters, (long long *)cu.getLongForceBuffer().getDevicePointer(),
                          (const long long *)cu1.getLongForceBuffer().getDevicePointer(),
                          (const long long *)cu2.getLongForceBuffer().getDevicePointer(), State1Energy, State2Energy, includeForces,
                          includeEnergy, cu.getCurrentStream());
}

void B200CalcATMMetaForceKernel::copyParametersToContext(ContextImpl &, const ATMMetaForce &force) {
    ContextSelector selector(cu);
    if (force.getNumParticles() == 0) return;
    kernel.copyParametersToContext(force, cu.getCurrentStream());
}
