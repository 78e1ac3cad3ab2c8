// openmm_standin_cuda.h -- stand-in for the slice of OpenMM's CUDA platform that a plugin kernel touches, for builds
// WITHOUT OpenMM: CudaArray (getDevicePointer / upload / download), CudaContext (getPosq, getPosqCorrection,
// getLongForceBuffer, getAtomIndex, getNumAtoms, getPaddedNumAtoms, getUseMixedPrecision, getUseDoublePrecision,
// getDeviceIndex, getCurrentStream, addReorderListener, setAsCurrent), ContextSelector and CudaPlatform with its
// PlatformData::contexts.  Method names follow OpenMM's documented plugin-facing API; the implementation
// (src/openmm_standin_cuda.cpp) is a few cudaMalloc / cudaMemcpy calls and is NOT copied from OpenMM.  Layout facts relied on
// (SURVEY.md appendix B): posq = float4 (x, y, z, q) padded to a multiple of 32; mixed precision keeps a float4 correction
// array; the long force buffer is int64[3][padded] with a 2^32 scale; atoms may be re-sorted, getAtomIndex()[slot] = atom.
//
// The platform also registers a "CalcNonbondedForce" kernel so that the inner contexts of ATMMetaForceImpl have
// something to evaluate: it plays the part of OpenMM's own CUDA NonbondedForce (direct space + PME reciprocal space +
// dispersion correction of ONE coordinate set) and is itself computed by libatm_b200 on a zero-displacement handle.
#ifndef ATM_OPENMM_STANDIN_CUDA_H_
#define ATM_OPENMM_STANDIN_CUDA_H_

#include <memory>
#include <string>
#include <vector>

#include "openmm_standin_context.h"

namespace OpenMM {

class CudaArray {
public:
    CudaArray() : ptr(nullptr), bytes(0) {}
    ~CudaArray();
    CudaArray(const CudaArray &) = delete;
    CudaArray &operator=(const CudaArray &) = delete;
    void initialize(size_t numBytes, const std::string &name);
    bool isInitialized() const { return ptr != nullptr; }
    void *getDevicePointer() { return ptr; }
    size_t getSize() const { return bytes; }
    void upload(const void *data);      // blocking, whole array
    void download(void *data) const;    // blocking, whole array
    void zero(void *stream);

private:
    void *ptr;
    size_t bytes;
    std::string name;
};

class CudaContext {
public:
    class ReorderListener {
    public:
        virtual ~ReorderListener() {}
        virtual void execute() = 0;
    };
    /** linked != NULL: an inner context sharing the device, the stream and the atom order of `linked`. */
    CudaContext(int numAtoms, int deviceIndex, bool mixedPrecision, bool doublePrecision, CudaContext *linked);
    ~CudaContext();
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    int getDeviceIndex() const { return deviceIndex; }
    bool getUseMixedPrecision() const { return mixed; }
    bool getUseDoublePrecision() const { return dbl; }
    CudaArray &getPosq() { return posq; }
    CudaArray &getPosqCorrection() { return posqCorrection; }
    CudaArray &getLongForceBuffer() { return force; }
    const std::vector<int> &getAtomIndex() const { return atomIndex; }
    void *getCurrentStream() { return stream; }
    void setAsCurrent();
    /** Takes ownership, as OpenMM does. */
    void addReorderListener(ReorderListener *listener) { listeners.emplace_back(listener); }
    /** What OpenMM's CudaContext::reorderAtoms() does when its neighbour list is rebuilt, with the new order given by
     *  the caller: posq (and the correction) are permuted on the device, atomIndex is replaced, the linked contexts
     *  follow, and every listener runs.  order[slot] = atom. */
    void reorderAtoms(const std::vector<int> &order);
    void synchronize();

private:
    int numAtoms, paddedNumAtoms, deviceIndex;
    bool mixed, dbl, ownsStream;
    void *stream;
    CudaArray posq, posqCorrection, force;
    std::vector<int> atomIndex;
    std::vector<std::unique_ptr<ReorderListener>> listeners;
    std::vector<CudaContext *> linkedContexts;
    CudaContext *parent;
};

/** RAII "make this context current" (OpenMM 7.7+ ContextSelector). */
class ContextSelector {
public:
    explicit ContextSelector(CudaContext &cu) { cu.setAsCurrent(); }
};

class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        std::vector<CudaContext *> contexts;   // one per device; the plugin uses contexts[0]
        ~PlatformData() { for (CudaContext *c : contexts) delete c; }
    };
    CudaPlatform();
    const std::string &getName() const override {
        static const std::string name = "CUDA";
        return name;
    }
    /** Property names of OpenMM's CUDA platform this stand-in honours: "DeviceIndex", "Precision" (single|mixed|double). */
    void contextCreated(ContextImpl &context, const std::map<std::string, std::string> &properties) const override;
    void linkedContextCreated(ContextImpl &context, ContextImpl &original) const override;
    void contextDestroyed(ContextImpl &context) const override;
    void setPositions(ContextImpl &context, const std::vector<Vec3> &positions) const override;
    void getPositions(const ContextImpl &context, std::vector<Vec3> &positions) const override;
    void beginComputation(ContextImpl &context, bool includeForces, bool includeEnergy, int groups) const override;
    double finishComputation(ContextImpl &context, bool includeForces, bool includeEnergy, int groups) const override;
    void getForces(ContextImpl &context, std::vector<Vec3> &forces) const override;
    void addForces(ContextImpl &context, const std::vector<Vec3> &forces) const override;
    static CudaContext &cudaContext(ContextImpl &context) {
        return *static_cast<PlatformData *>(context.getPlatformData())->contexts[0];
    }
};

}  // namespace OpenMM

#endif
