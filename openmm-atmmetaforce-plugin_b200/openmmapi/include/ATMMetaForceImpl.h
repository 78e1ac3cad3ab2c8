// ATMMetaForceImpl.h -- the per-Context implementation object of ATMMetaForce, with the surface of the reference's
// ATMMetaForceImpl (ref: openmmapi/include/internal/ATMMetaForceImpl.h:24-60): initialize, calcForcesAndEnergy,
// getDefaultParameters, getKernelNames, updateParametersInContext, getPerturbationEnergy.
//
// Build WITHOUT OpenMM (this image): the Impl is the whole orchestration of the Tier-2 path.  Where the reference
// creates two inner Contexts and evaluates them one after the other (ref: openmmapi/src/ATMMetaForceImpl.cpp:90-128),
// this Impl describes the variable-group NonbondedForce to the back-end once (atm_nb_setup) and then, per evaluation,
// hands the Context's host positions to atm_host_pipeline_step -- H2D, [rebuild | prune], copy-state/pack, the two-state
// direct-space kernel, the device scalar stage, the merge, D2H of forces and energy record -- one CUDA-graph launch.
// The NonbondedForce is evaluated as a whole -- direct space, PME reciprocal space of both states and the dispersion
// correction, as the reference's inner contexts do -- unless its reciprocal space was moved to a non-variable group.
// Any OTHER Force in a variable force group (OpenMM::HostEvaluatedForce in this build, e.g. HarmonicBondForce) is
// evaluated at x and at x + d as well and enters the same step through atm_host_io.force_state{1,2}_ext_host /
// energy_ext_host -- the generic hook for what the reference's inner contexts evaluate besides the NonbondedForce.
//
// When the Context's Platform has a "CalcATMMetaForce" kernel registered (libATMMetaForcePluginCUDA.so on the "CUDA"
// platform), the Impl instead runs the REFERENCE's orchestration, unchanged in structure: every non-ATM Force of the
// System is cloned into two inner Systems (ref: copysystem, ATMMetaForceImpl.cpp:33-67), two linked inner Contexts
// evaluate whatever sits in the variable force groups at x and at x + d, and the kernel seam does copyState and the
// hybrid merge (ref: :90-128).  That is the generic path: any Force with an Impl on that platform can be variable.
#ifndef ATMMETAFORCE_IMPL_H_
#define ATMMETAFORCE_IMPL_H_

#include <map>
#include <string>
#include <vector>

#include "ATMMetaForce.h"
#include "atm_b200.h"
#ifndef ATM_HAVE_OPENMM
#include "openmm_standin_context.h"
#endif
#include <memory>

namespace ATMMetaForcePlugin {

#ifndef ATM_HAVE_OPENMM

class ATMMetaForceImpl : public OpenMM::ForceImpl {
public:
    explicit ATMMetaForceImpl(const ATMMetaForce &owner);
    ~ATMMetaForceImpl();
    ATMMetaForceImpl(const ATMMetaForceImpl &) = delete;
    ATMMetaForceImpl &operator=(const ATMMetaForceImpl &) = delete;

    /** ref: ATMMetaForceImpl::initialize (:69-88): the variable-force-group mask (the ATM group itself cannot be
     *  variable), and which Forces of the System the two states evaluate.  No device work yet. */
    void initialize(OpenMM::ContextImpl &context) override;
    const ATMMetaForce &getOwner() const { return owner; }
    void updateContextState(OpenMM::ContextImpl &, bool & /*forcesInvalid*/) {}
    /** ref: ATMMetaForceImpl::calcForcesAndEnergy (:90-128).  Adds the ATM force to context.getForces(); returns
     *  e0 + W(u_sc) when includeEnergy, else 0; 0 without any work when the ATM group is not in `groups`. */
    double calcForcesAndEnergy(OpenMM::ContextImpl &context, bool includeForces, bool includeEnergy, int groups) override;
    /** ref: :130-142 */
    std::map<std::string, double> getDefaultParameters() override;
    std::vector<std::string> getKernelNames() override;
    /** ref: :150-152 -> kernel copyParametersToContext (CommonATMMetaForceKernels.cpp:229-251). */
    void updateParametersInContext(OpenMM::ContextImpl &context);
    double getPerturbationEnergy() const { return PerturbationEnergy; }

    // knobs of this back-end (no reference counterpart)
    void setPairListSkins(double inner_nm, double outer_nm);
    void setDevice(int ordinal) { device = ordinal; }
    int getVariableForceGroupsMask() const { return variable_force_groups_mask; }
    /** The full energy record of the last evaluation (atm_energy_slot order). */
    const std::vector<double> &getEnergyRecord() const { return energyRecord; }
    /** True when the platform's CalcATMMetaForce kernel and two inner Contexts are used (the reference's path). */
    bool usesPlatformKernel() const { return (bool)kernel; }
    OpenMM::Context *getInnerContext(int state) { return state == 1 ? innerContext1.get() : innerContext2.get(); }

private:
    void createBackend(OpenMM::ContextImpl &context);
    void releaseBackend();
    /** ref: ATMMetaForceImpl::copysystem (:33-67). */
    void copysystem(const OpenMM::System &system, OpenMM::System &innerSystem);
    double calcWithPlatformKernel(OpenMM::ContextImpl &context, bool includeForces, bool includeEnergy, int groups);

    // the reference's orchestration (platform kernel + two linked inner contexts)
    OpenMM::Kernel kernel;
    OpenMM::System innerSystem1, innerSystem2;
    OpenMM::VerletIntegrator innerIntegrator1, innerIntegrator2;
    std::unique_ptr<OpenMM::Context> innerContext1, innerContext2;
    bool hasInitializedInnerContexts;

    const ATMMetaForce &owner;
    const OpenMM::NonbondedForce *nonbonded;
    double PerturbationEnergy;
    int variable_force_groups_mask;
    int device;
    double skin, skinOuter;
    // back-end objects (created at the first evaluation, so that a Context can be built and configured on any host)
    atm_handle *handle;
    atm_host_pipeline *pipeline;
    void *stream;
    float *posqHost;       // pinned [P][4]
    int64_t *forceHost;    // pinned [3P]
    double *energyHost;    // pinned [ATM_NUM_ENERGY_SLOTS]
    // the other Forces of the variable force groups (host-evaluated in this build) and their per-state results
    std::vector<const OpenMM::HostEvaluatedForce *> hostForces;
    int64_t *extHost[2];   // pinned [3P] each: forces of hostForces at the state-1 / state-2 coordinates, 2^32 fixed point
    double *energyExtHost; // pinned [2]
    std::vector<double> displacements;   // [N][3], refreshed with the device table
    int paddedNumAtoms;
    bool displacementsDirty, reciprocalOn;
    unsigned long boxVersionSeen;
    std::vector<OpenMM::Vec3> refRebuild, refPrune;
    std::vector<double> energyRecord;
};

#endif  // !ATM_HAVE_OPENMM

}  // namespace ATMMetaForcePlugin

#endif
