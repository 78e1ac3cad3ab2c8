// ATMMetaForceB200Kernel.h -- host-side kernel object of the single Blackwell back-end.
//
// It has the five operations of the reference's kernel seam CalcATMMetaForceKernel
// (ref: openmmapi/include/ATMMetaForceKernels.h:15-59): initialize, copyState, execute, copyParametersToContext,
// getPerturbationEnergy -- with OpenMM's ContextImpl arguments replaced by the raw device buffers those contexts own
// (posq, posqCorrection, long force buffers, atom index; ref: platforms/common/src/CommonATMMetaForceKernels.cpp:100,
// 128-135,142-145), because OpenMM itself is not available in this build.  INTEGRATION.md shows the ten-line adapter
// that derives it from OpenMM::KernelImpl when OpenMM is present.  All numerics live below the C ABI (atm_b200.h).
#ifndef ATMMETAFORCE_B200_KERNEL_H_
#define ATMMETAFORCE_B200_KERNEL_H_

#include <map>
#include <string>
#include <vector>

#include "ATMMetaForce.h"
#include "atm_b200.h"

namespace ATMMetaForcePlugin {

class ATMMetaForceB200Kernel {
public:
    static std::string Name() { return "CalcATMMetaForce"; }
    ATMMetaForceB200Kernel() : handle(nullptr), numParticles(0), perturbationEnergy(0.0) {}
    ~ATMMetaForceB200Kernel();
    ATMMetaForceB200Kernel(const ATMMetaForceB200Kernel &) = delete;
    ATMMetaForceB200Kernel &operator=(const ATMMetaForceB200Kernel &) = delete;

    /** ref: CommonCalcATMMetaForceKernel::initialize (CommonATMMetaForceKernels.cpp:78-109).
     *  atomIndex = cc.getAtomIndex() (slot -> atom), may be empty for the identity order. */
    void initialize(const ATMMetaForce &force, int paddedNumAtoms, atm_precision precision,
                    const std::vector<int> &atomIndex, int device = -1);
    /** ref: ReorderListener::execute (:53-72) -- call whenever OpenMM re-sorts the atoms. */
    void atomsReordered(const ATMMetaForce &force, const std::vector<int> &atomIndex, void *stream = nullptr);
    /** ref: copyState (:206-212): device part only (the host-side box/time/parameter mirroring needs OpenMM). */
    void copyState(const void *posq, const void *posqCorrection, void *posq1, void *posq1Correction, void *posq2,
                   void *posq2Correction, void *stream = nullptr);
    /** ref: execute (:154-204).  `parameters` holds the Context's global parameters by name. */
    double execute(const std::map<std::string, double> &parameters, long long *force, const long long *forceState1,
                   const long long *forceState2, double State1Energy, double State2Energy, bool includeForces,
                   bool includeEnergy, void *stream = nullptr);
    /** ref: copyParametersToContext (:229-251). */
    void copyParametersToContext(const ATMMetaForce &force, void *stream = nullptr);
    double getPerturbationEnergy() const { return perturbationEnergy; }
    /** ref: ATMMetaForceImpl::getDefaultParameters (openmmapi/src/ATMMetaForceImpl.cpp:130-142). */
    static std::map<std::string, double> getDefaultParameters(const ATMMetaForce &force);
    atm_handle *getHandle() { return handle; }

private:
    atm_handle *handle;
    int numParticles;
    std::vector<int> atomIndex;
    double perturbationEnergy;
};

/** Variable-force-group mask exactly as ATMMetaForceImpl::initialize builds it (ref: ATMMetaForceImpl.cpp:75-81),
 *  including the rule that the ATM force group itself cannot be variable (throws OpenMMException). */
int variableForceGroupsMask(const ATMMetaForce &force);

}  // namespace ATMMetaForcePlugin
#endif
