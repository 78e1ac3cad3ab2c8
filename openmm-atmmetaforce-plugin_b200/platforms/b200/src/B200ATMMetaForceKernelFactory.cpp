// B200ATMMetaForceKernelFactory.cpp -- plugin registration of the Blackwell back-end on OpenMM's "CUDA" platform.
// Exports the same three C symbols as the reference's CUDA plugin library
// (ref: platforms/cuda/src/CudaATMMetaForceKernelFactory.cpp:14-43): registerPlatforms, registerKernelFactories (what
// Platform::loadPluginLibrary calls) and registerATMMetaForceCudaKernelFactories (what a statically linked host calls;
// it registers the CUDA platform itself when it is not there yet).
#include "B200ATMMetaForceKernelFactory.h"

#include <exception>

#include "B200ATMMetaForceKernels.h"

using namespace ATMMetaForcePlugin;
using namespace OpenMM;

#if defined(_WIN32)
#define ATM_PLUGIN_EXPORT __declspec(dllexport)
#else
#define ATM_PLUGIN_EXPORT __attribute__((visibility("default")))
#endif

extern "C" ATM_PLUGIN_EXPORT void registerPlatforms() {}

extern "C" ATM_PLUGIN_EXPORT void registerKernelFactories() {
    try {
        Platform &platform = Platform::getPlatformByName("CUDA");
        platform.registerKernelFactory(CalcATMMetaForceKernel::Name(), new B200ATMMetaForceKernelFactory());
    } catch (const std::exception &) {
        // no CUDA platform in this process: nothing to register on
    }
}

extern "C" ATM_PLUGIN_EXPORT void registerATMMetaForceCudaKernelFactories() {
    try {
        Platform::getPlatformByName("CUDA");
    } catch (...) {
        Platform::registerPlatform(new CudaPlatform());
    }
    registerKernelFactories();
}

KernelImpl *B200ATMMetaForceKernelFactory::createKernelImpl(std::string name, const Platform &platform, ContextImpl &context) const {
    CudaContext &cu = *static_cast<CudaPlatform::PlatformData *>(context.getPlatformData())->contexts[0];
    if (name == CalcATMMetaForceKernel::Name()) return new B200CalcATMMetaForceKernel(name, platform, cu);
    throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
}
