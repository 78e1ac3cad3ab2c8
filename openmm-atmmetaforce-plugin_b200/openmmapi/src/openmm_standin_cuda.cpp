// openmm_standin_cuda.cpp -- implementation of the CUDA-platform stand-in (openmm_standin_cuda.h): device arrays through
// the CUDA runtime, and the platform's own "CalcNonbondedForce" kernel (the part OpenMM's CUDA NonbondedForce plays in a
// real installation), evaluated by libatm_b200 on a zero-displacement handle.  Not compiled with -DATM_HAVE_OPENMM.
#ifndef ATM_HAVE_OPENMM
#include "openmm_standin_cuda.h"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "atm_b200.h"

namespace OpenMM {

namespace {
void cudaCheck(cudaError_t err, const char *what) {
    if (err != cudaSuccess) throw OpenMMException(std::string("CUDA stand-in: ") + what + ": " + cudaGetErrorString(err));
}
void atmCheck(int rc, const char *what) {
    if (rc != ATM_OK) throw OpenMMException(std::string(what) + ": " + atm_last_error());
}
}  // namespace

// ---- CudaArray ------------------------------------------------------------------------------------------------------
CudaArray::~CudaArray() {
    if (ptr) cudaFree(ptr);
}
void CudaArray::initialize(size_t numBytes, const std::string &n) {
    if (ptr) throw OpenMMException("CudaArray has already been initialized");
    name = n;
    bytes = numBytes;
    cudaCheck(cudaMalloc(&ptr, std::max<size_t>(numBytes, 1)), ("allocating " + n).c_str());
    cudaCheck(cudaMemset(ptr, 0, numBytes), ("clearing " + n).c_str());
}
void CudaArray::upload(const void *data) { cudaCheck(cudaMemcpy(ptr, data, bytes, cudaMemcpyHostToDevice), ("uploading " + name).c_str()); }
void CudaArray::download(void *data) const { cudaCheck(cudaMemcpy(data, ptr, bytes, cudaMemcpyDeviceToHost), ("downloading " + name).c_str()); }
void CudaArray::zero(void *stream) { cudaCheck(cudaMemsetAsync(ptr, 0, bytes, (cudaStream_t)stream), ("clearing " + name).c_str()); }

// ---- CudaContext ----------------------------------------------------------------------------------------------------
CudaContext::CudaContext(int numAtoms, int deviceIndex, bool mixedPrecision, bool doublePrecision, CudaContext *linked)
    : numAtoms(numAtoms), paddedNumAtoms(32 * ((numAtoms + 31) / 32)), deviceIndex(deviceIndex), mixed(mixedPrecision),
      dbl(doublePrecision), ownsStream(false), stream(nullptr), parent(linked) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw OpenMMException("No compatible CUDA device is available");
    if (this->deviceIndex < 0) this->deviceIndex = linked ? linked->deviceIndex : 0;
    cudaCheck(cudaSetDevice(this->deviceIndex), "cudaSetDevice");
    if (linked) {
        stream = linked->stream;
        atomIndex = linked->atomIndex;
        linked->linkedContexts.push_back(this);
    } else {
        cudaStream_t s;
        cudaCheck(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
        stream = s;
        ownsStream = true;
        atomIndex.resize(numAtoms);
        for (int i = 0; i < numAtoms; i++) atomIndex[i] = i;
    }
    const size_t elem = dbl ? 4 * sizeof(double) : 4 * sizeof(float);
    posq.initialize(elem * paddedNumAtoms, "posq");
    if (mixed) posqCorrection.initialize(4 * sizeof(float) * paddedNumAtoms, "posqCorrection");
    force.initialize(3 * sizeof(long long) * (size_t)paddedNumAtoms, "longForceBuffer");
}

CudaContext::~CudaContext() {
    if (parent) {
        auto &v = parent->linkedContexts;
        v.erase(std::remove(v.begin(), v.end(), this), v.end());
    }
    for (CudaContext *c : linkedContexts) c->parent = nullptr;
    listeners.clear();
    if (ownsStream && stream) {
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaStreamDestroy((cudaStream_t)stream);
    }
}

void CudaContext::setAsCurrent() { cudaSetDevice(deviceIndex); }
void CudaContext::synchronize() { cudaCheck(cudaStreamSynchronize((cudaStream_t)stream), "cudaStreamSynchronize"); }

void CudaContext::reorderAtoms(const std::vector<int> &order) {
    if ((int)order.size() != numAtoms) throw OpenMMException("reorderAtoms: wrong length");
    synchronize();
    // permute posq (+ correction) on the host: this is a test facility, not a hot path
    std::vector<int> slotOfAtom(numAtoms);
    for (int s = 0; s < numAtoms; s++) slotOfAtom[atomIndex[s]] = s;
    auto permute = [&](CudaArray &a, size_t elem) {
        if (!a.isInitialized()) return;
        std::vector<char> oldv(a.getSize()), newv(a.getSize(), 0);
        a.download(oldv.data());
        for (int s = 0; s < numAtoms; s++) std::memcpy(&newv[elem * s], &oldv[elem * slotOfAtom[order[s]]], elem);
        a.upload(newv.data());
    };
    permute(posq, dbl ? 4 * sizeof(double) : 4 * sizeof(float));
    permute(posqCorrection, 4 * sizeof(float));
    atomIndex = order;
    for (auto &l : listeners) l->execute();
    for (CudaContext *c : linkedContexts) {   // inner contexts keep the outer order (their posq is overwritten by copyState)
        c->atomIndex = order;
        for (auto &l : c->listeners) l->execute();
    }
}

// ---- the platform's NonbondedForce kernel ---------------------------------------------------------------------------------
namespace {
class StandinCudaCalcNonbondedForceKernel : public CalcNonbondedForceKernel {
public:
    StandinCudaCalcNonbondedForceKernel(std::string name, const Platform &platform, CudaContext &cu)
        : CalcNonbondedForceKernel(name, platform), cu(cu), handle(nullptr), boxVersion(~0ul), hasPme(false) {}
    ~StandinCudaCalcNonbondedForceKernel() {
        if (handle) atm_destroy(handle);
    }
    void initialize(const System &system, const NonbondedForce &force) override {
        ContextSelector selector(cu);
        if (cu.getUseDoublePrecision()) throw OpenMMException("CUDA stand-in: NonbondedForce supports single and mixed precision");
        if (force.getNonbondedMethod() != NonbondedForce::PME && force.getNonbondedMethod() != NonbondedForce::Ewald)
            throw OpenMMException("CUDA stand-in: NonbondedForce supports periodic Ewald / PME only");
        const int n = system.getNumParticles();
        atm_config cfg;
        cfg.num_particles = n;
        cfg.padded_num_particles = cu.getPaddedNumAtoms();
        cfg.precision = cu.getUseMixedPrecision() ? ATM_PREC_MIXED : ATM_PREC_SINGLE;
        cfg.num_replicas = 1;
        cfg.device = cu.getDeviceIndex();
        atmCheck(atm_create(&cfg, &handle), "CUDA stand-in: NonbondedForce back-end");
        zeros.assign(3 * (size_t)n, 0.0);
        atmCheck(atm_set_displacements(handle, cu.getAtomIndex().data(), zeros.data(), cu.getCurrentStream()), "CUDA stand-in: atom order");
        const double p[ATM_NUM_PARAMS] = {0, 0, 0, 0, 0, 1e6, 5e5, 0.0625, 1};
        atmCheck(atm_set_parameters(handle, 0, p), "CUDA stand-in: parameters");
        q.resize(n); sig.resize(n); eps.resize(n);
        for (int i = 0; i < n; i++) force.getParticleParameters(i, q[i], sig[i], eps[i]);
        for (int e = 0; e < force.getNumExceptions(); e++) {
            int a, b;
            double cp, s, ep;
            force.getExceptionParameters(e, a, b, cp, s, ep);
            excl.push_back(a); excl.push_back(b);
            if (cp != 0.0 || ep != 0.0) {
                excPairs.push_back(a); excPairs.push_back(b);
                excParams.push_back(cp); excParams.push_back(s); excParams.push_back(ep);
            }
        }
        cutoff = force.getCutoffDistance();
        tolerance = force.getEwaldErrorTolerance();
        alpha = std::sqrt(-std::log(2.0 * tolerance)) / cutoff;
        hasPme = true;   // both Ewald and PME are evaluated with the smooth-PME mesh of OpenMM's rule
        dispersion = force.getUseDispersionCorrection();
        cu.addReorderListener(new Reorder(*this));
    }
    double execute(ContextImpl &context, bool, bool includeEnergy, bool includeDirect, bool includeReciprocal) override {
        ContextSelector selector(cu);
        if (!(includeDirect && includeReciprocal))
            throw OpenMMException("CUDA stand-in: direct and reciprocal space must be in the same force group");
        void *stream = cu.getCurrentStream();
        if (boxVersion != context.getBoxVersion()) {
            Vec3 box[3];
            context.getPeriodicBoxVectors(box[0], box[1], box[2]);
            const double b[9] = {box[0][0], box[0][1], box[0][2], box[1][0], box[1][1], box[1][2], box[2][0], box[2][1], box[2][2]};
            atmCheck(atm_set_box(handle, -1, b), "CUDA stand-in: periodic box");
            if (boxVersion == ~0ul) {
                atm_nonbonded_desc desc;
                std::memset(&desc, 0, sizeof(desc));
                desc.charge = q.data(); desc.sigma = sig.data(); desc.epsilon = eps.data();
                desc.num_exclusions = (int32_t)(excl.size() / 2); desc.exclusions = excl.data();
                desc.num_exceptions = (int32_t)(excPairs.size() / 2); desc.exception_pairs = excPairs.data(); desc.exception_params = excParams.data();
                desc.cutoff = cutoff; desc.ewald_alpha = alpha;
                desc.skin = 0.02; desc.skin_outer = 0.02;   // the lists are rebuilt at every evaluation
                atmCheck(atm_nb_setup(handle, &desc, stream), "CUDA stand-in: NonbondedForce set-up");
                atmCheck(atm_nb_set_dispersion_correction(handle, dispersion ? 1 : 0), "CUDA stand-in: dispersion correction");
            }
            if (hasPme) {
                int grid[3];
                pmeGridDimensions(alpha, tolerance, box, grid);
                atmCheck(atm_pme_setup(handle, grid[0], grid[1], grid[2], 5), "CUDA stand-in: PME mesh");
            }
            boxVersion = context.getBoxVersion();
        }
        void *posq = cu.getPosq().getDevicePointer();
        atmCheck(atm_nb_rebuild(handle, posq, stream), "CUDA stand-in: neighbour list");
        atm_step_io io;
        std::memset(&io, 0, sizeof(io));
        io.posq = posq;
        io.posq_corr = cu.getUseMixedPrecision() ? cu.getPosqCorrection().getDevicePointer() : nullptr;
        io.force = (int64_t *)cu.getLongForceBuffer().getDevicePointer();
        io.include_energy = 1;
        atmCheck(atm_step(handle, &io, stream), "CUDA stand-in: NonbondedForce evaluation");
        double rec[ATM_NUM_ENERGY_SLOTS];
        atmCheck(atm_get_energies(handle, rec, stream), "CUDA stand-in: energy download");   // blocking, like OpenMM's energy read-back
        return includeEnergy ? rec[ATM_E_U1] : 0.0;
    }

private:
    struct Reorder : CudaContext::ReorderListener {
        explicit Reorder(StandinCudaCalcNonbondedForceKernel &k) : k(k) {}
        void execute() override {
            atmCheck(atm_set_displacements(k.handle, k.cu.getAtomIndex().data(), k.zeros.data(), k.cu.getCurrentStream()), "CUDA stand-in: atom order");
        }
        StandinCudaCalcNonbondedForceKernel &k;
    };
    CudaContext &cu;
    atm_handle *handle;
    unsigned long boxVersion;
    bool hasPme, dispersion;
    double cutoff, tolerance, alpha;
    std::vector<double> q, sig, eps, excParams, zeros;
    std::vector<int32_t> excl, excPairs;
};

class StandinCudaKernelFactory : public KernelFactory {
public:
    KernelImpl *createKernelImpl(std::string name, const Platform &platform, ContextImpl &context) const override {
        CudaContext &cu = *static_cast<CudaPlatform::PlatformData *>(context.getPlatformData())->contexts[0];
        if (name == CalcNonbondedForceKernel::Name()) return new StandinCudaCalcNonbondedForceKernel(name, platform, cu);
        throw OpenMMException("Tried to create kernel with illegal kernel name '" + name + "'");
    }
};
}  // namespace

// ---- CudaPlatform ---------------------------------------------------------------------------------------------------
CudaPlatform::CudaPlatform() { registerKernelFactory(CalcNonbondedForceKernel::Name(), new StandinCudaKernelFactory()); }

void CudaPlatform::contextCreated(ContextImpl &context, const std::map<std::string, std::string> &properties) const {
    int device = -1;
    std::string precision = "mixed";
    auto it = properties.find("DeviceIndex");
    if (it != properties.end()) device = std::atoi(it->second.c_str());
    it = properties.find("Precision");
    if (it != properties.end()) precision = it->second;
    if (precision != "single" && precision != "mixed" && precision != "double")
        throw OpenMMException("Illegal value for Precision: " + precision);
    auto *data = new PlatformData();
    try {
        data->contexts.push_back(new CudaContext(context.getSystem().getNumParticles(), device, precision == "mixed", precision == "double", nullptr));
    } catch (...) {
        delete data;
        throw;
    }
    context.setPlatformData(data);
}

void CudaPlatform::linkedContextCreated(ContextImpl &context, ContextImpl &original) const {
    CudaContext &cu = cudaContext(original);
    auto *data = new PlatformData();
    try {
        data->contexts.push_back(new CudaContext(context.getSystem().getNumParticles(), cu.getDeviceIndex(), cu.getUseMixedPrecision(),
                                                 cu.getUseDoublePrecision(), &cu));
    } catch (...) {
        delete data;
        throw;
    }
    context.setPlatformData(data);
}

void CudaPlatform::contextDestroyed(ContextImpl &context) const {
    delete static_cast<PlatformData *>(context.getPlatformData());
    context.setPlatformData(nullptr);
}

void CudaPlatform::setPositions(ContextImpl &context, const std::vector<Vec3> &positions) const {
    CudaContext &cu = cudaContext(context);
    ContextSelector selector(cu);
    cu.synchronize();
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    const std::vector<int> &order = cu.getAtomIndex();
    if (cu.getUseDoublePrecision()) {
        std::vector<double> p(4 * (size_t)P, 0.0);
        for (int s = 0; s < n; s++)
            for (int c = 0; c < 3; c++) p[4 * (size_t)s + c] = positions[order[s]][c];
        cu.getPosq().upload(p.data());
        return;
    }
    std::vector<float> p(4 * (size_t)P, 0.f), corr(4 * (size_t)P, 0.f);
    for (int s = 0; s < n; s++)
        for (int c = 0; c < 3; c++) {
            const double x = positions[order[s]][c];
            p[4 * (size_t)s + c] = (float)x;
            corr[4 * (size_t)s + c] = (float)(x - (double)(float)x);
        }
    cu.getPosq().upload(p.data());
    if (cu.getUseMixedPrecision()) cu.getPosqCorrection().upload(corr.data());
}

// What the device holds (an inner context's coordinates are written there by the ATM kernel's copyState, never by
// setPositions): posq (+ the correction in mixed precision), slot order -> atom order.
void CudaPlatform::getPositions(const ContextImpl &context, std::vector<Vec3> &positions) const {
    CudaContext &cu = cudaContext(const_cast<ContextImpl &>(context));
    ContextSelector selector(cu);
    cu.synchronize();
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    const std::vector<int> &order = cu.getAtomIndex();
    positions.assign(n, Vec3());
    if (cu.getUseDoublePrecision()) {
        std::vector<double> p(4 * (size_t)P);
        cu.getPosq().download(p.data());
        for (int s = 0; s < n; s++)
            for (int c = 0; c < 3; c++) positions[order[s]][c] = p[4 * (size_t)s + c];
        return;
    }
    std::vector<float> p(4 * (size_t)P), corr(4 * (size_t)P, 0.f);
    cu.getPosq().download(p.data());
    if (cu.getUseMixedPrecision()) cu.getPosqCorrection().download(corr.data());
    for (int s = 0; s < n; s++)
        for (int c = 0; c < 3; c++) positions[order[s]][c] = (double)p[4 * (size_t)s + c] + (double)corr[4 * (size_t)s + c];
}

// Host-evaluated forces enter the long force buffer the way every OpenMM kernel's do: 2^32 fixed point, slot order.
void CudaPlatform::addForces(ContextImpl &context, const std::vector<Vec3> &forces) const {
    CudaContext &cu = cudaContext(context);
    ContextSelector selector(cu);
    cu.synchronize();
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    std::vector<long long> f(3 * (size_t)P);
    cu.getLongForceBuffer().download(f.data());
    const std::vector<int> &order = cu.getAtomIndex();
    for (int s = 0; s < n; s++)
        for (int c = 0; c < 3; c++) f[(size_t)c * P + s] += (long long)std::llrint(forces[order[s]][c] * 4294967296.0);
    cu.getLongForceBuffer().upload(f.data());
}

void CudaPlatform::beginComputation(ContextImpl &context, bool, bool, int) const {
    CudaContext &cu = cudaContext(context);
    ContextSelector selector(cu);
    cu.getLongForceBuffer().zero(cu.getCurrentStream());
}

double CudaPlatform::finishComputation(ContextImpl &context, bool, bool, int) const {
    cudaContext(context).synchronize();
    return 0.0;
}

void CudaPlatform::getForces(ContextImpl &context, std::vector<Vec3> &forces) const {
    CudaContext &cu = cudaContext(context);
    ContextSelector selector(cu);
    cu.synchronize();
    const int P = cu.getPaddedNumAtoms(), n = cu.getNumAtoms();
    std::vector<long long> f(3 * (size_t)P);
    cu.getLongForceBuffer().download(f.data());
    forces.assign(n, Vec3());
    const std::vector<int> &order = cu.getAtomIndex();
    for (int s = 0; s < n; s++)
        for (int c = 0; c < 3; c++) forces[order[s]][c] = (double)f[(size_t)c * P + s] / 4294967296.0;
}

}  // namespace OpenMM
#endif
