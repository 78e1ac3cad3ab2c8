// pybind11 binding of the C++ facade (plays the role of the reference's SWIG wrapper, python/atmmetaforceplugin.i,
// which cannot be generated here: SWIG and OpenMM's swig headers are absent).  Same method names and argument order;
// getParticleParameters returns the tuple (particle, dx, dy, dz) like the SWIG OUTPUT typemaps (.i:79-87);
// std::exception is mapped to a Python exception like the %exception block (.i:54-61).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>

#include "ATMMetaForce.h"
#include "ATMMetaForceB200Kernel.h"
#include "ATMMetaForceImpl.h"
#include "ATMMetaForceProxy.h"
#include "openmm_standin_cuda.h"

namespace py = pybind11;
using namespace ATMMetaForcePlugin;

static void *ptr(uintptr_t p) { return reinterpret_cast<void *>(p); }

static OpenMM::Vec3 vec3(const std::vector<double> &v) {
    if (v.size() != 3) throw OpenMM::OpenMMException("expected a vector of three numbers");
    return OpenMM::Vec3(v[0], v[1], v[2]);
}

PYBIND11_MODULE(_atmmetaforce_core, m) {
    m.doc() = "C++ facade of the Blackwell ATM Meta-Force back-end";
    py::register_exception<OpenMM::OpenMMException>(m, "OpenMMException", PyExc_Exception);
    m.attr("ATMMETAFORCE_VERSION") = ATMMETAFORCE_VERSION;

    py::class_<ATMMetaForce>(m, "ATMMetaForce")
        .def(py::init<double, double, double, double, double, double, double, double, double, const std::vector<int> &>(),
             py::arg("Lambda1"), py::arg("Lambda2"), py::arg("Alpha"), py::arg("U0"), py::arg("W0"), py::arg("Umax"),
             py::arg("Ubcore"), py::arg("Acore"), py::arg("direction"), py::arg("VariableForceGroups"))
        .def("getNumParticles", &ATMMetaForce::getNumParticles)
        .def("addParticle", &ATMMetaForce::addParticle)
        .def("setParticleParameters", &ATMMetaForce::setParticleParameters)
        .def("getParticleParameters",
             [](const ATMMetaForce &f, int index) {
                 int particle;
                 double dx, dy, dz;
                 f.getParticleParameters(index, particle, dx, dy, dz);
                 return py::make_tuple(particle, dx, dy, dz);
             })
        .def("usesPeriodicBoundaryConditions", &ATMMetaForce::usesPeriodicBoundaryConditions)
        .def("getForceGroup", &ATMMetaForce::getForceGroup)
        .def("setForceGroup", &ATMMetaForce::setForceGroup)
        .def("getName", &ATMMetaForce::getName)
        .def("setName", &ATMMetaForce::setName)
        .def_static("Lambda1", &ATMMetaForce::Lambda1)
        .def_static("Lambda2", &ATMMetaForce::Lambda2)
        .def_static("Alpha", &ATMMetaForce::Alpha)
        .def_static("U0", &ATMMetaForce::U0)
        .def_static("W0", &ATMMetaForce::W0)
        .def_static("Umax", &ATMMetaForce::Umax)
        .def_static("Ubcore", &ATMMetaForce::Ubcore)
        .def_static("Acore", &ATMMetaForce::Acore)
        .def_static("Direction", &ATMMetaForce::Direction)
        .def_static("Version", &ATMMetaForce::Version)
        .def("getDefaultLambda1", &ATMMetaForce::getDefaultLambda1)
        .def("getDefaultLambda2", &ATMMetaForce::getDefaultLambda2)
        .def("getDefaultAlpha", &ATMMetaForce::getDefaultAlpha)
        .def("getDefaultU0", &ATMMetaForce::getDefaultU0)
        .def("getDefaultW0", &ATMMetaForce::getDefaultW0)
        .def("getDefaultUmax", &ATMMetaForce::getDefaultUmax)
        .def("getDefaultUbcore", &ATMMetaForce::getDefaultUbcore)
        .def("getDefaultAcore", &ATMMetaForce::getDefaultAcore)
        .def("getDefaultDirection", &ATMMetaForce::getDefaultDirection)
        .def("getVariableForceGroups", &ATMMetaForce::getVariableForceGroups)
        .def("getPerturbationEnergy", &ATMMetaForce::getPerturbationEnergy, py::arg("context"))
        .def("updateParametersInContext", &ATMMetaForce::updateParametersInContext, py::arg("context"))
        .def("getDisplacementArray", &ATMMetaForce::getDisplacementArray)
        .def("getDefaultParameterArray", [](const ATMMetaForce &f) {
            std::vector<double> p(9);
            f.getDefaultParameters(p.data());
            return p;
        });

    m.def("serialize", [](const ATMMetaForce &f, const std::string &rootName) {
        std::stringstream ss;
        OpenMM::XmlSerializer::serialize<ATMMetaForce>(&f, rootName, ss, "ATMMetaForce");
        return ss.str();
    }, py::arg("force"), py::arg("rootName") = "Force");
    m.def("deserialize", [](const std::string &xml) {
        std::stringstream ss(xml);
        return std::unique_ptr<ATMMetaForce>(OpenMM::XmlSerializer::deserialize<ATMMetaForce>(ss));
    });
    m.def("variableForceGroupsMask", &variableForceGroupsMask);

    // ---- the OpenMM-free System / Context of this build (openmm_standin_context.h) and the C++ ATMMetaForceImpl behind it
    py::class_<OpenMM::System>(m, "System")
        .def(py::init<>())
        .def("addParticle", &OpenMM::System::addParticle, py::arg("mass"))
        .def("getNumParticles", &OpenMM::System::getNumParticles)
        .def("getNumForces", &OpenMM::System::getNumForces)
        .def("setDefaultPeriodicBoxVectors", [](OpenMM::System &s, const std::vector<double> &a, const std::vector<double> &b,
                                                const std::vector<double> &c) { s.setDefaultPeriodicBoxVectors(vec3(a), vec3(b), vec3(c)); })
        .def("addNonbondedForce", [](OpenMM::System &s, const std::vector<double> &charge, const std::vector<double> &sigma,
                                     const std::vector<double> &epsilon, const std::vector<int> &exceptionPairs,
                                     const std::vector<double> &exceptionParams, double cutoff, double ewaldTolerance, int group,
                                     int reciprocalSpaceForceGroup, bool useDispersionCorrection) {
            if (charge.size() != sigma.size() || charge.size() != epsilon.size() || exceptionPairs.size() % 2 != 0 ||
                exceptionParams.size() / 3 != exceptionPairs.size() / 2)
                throw OpenMM::OpenMMException("addNonbondedForce: inconsistent array lengths");
            auto *nb = new OpenMM::NonbondedForce();
            for (size_t i = 0; i < charge.size(); i++) nb->addParticle(charge[i], sigma[i], epsilon[i]);
            for (size_t e = 0; e < exceptionPairs.size() / 2; e++)
                nb->addException(exceptionPairs[2 * e], exceptionPairs[2 * e + 1], exceptionParams[3 * e], exceptionParams[3 * e + 1],
                                 exceptionParams[3 * e + 2]);
            nb->setNonbondedMethod(OpenMM::NonbondedForce::PME);
            nb->setCutoffDistance(cutoff);
            nb->setEwaldErrorTolerance(ewaldTolerance);
            nb->setForceGroup(group);
            nb->setReciprocalSpaceForceGroup(reciprocalSpaceForceGroup);
            nb->setUseDispersionCorrection(useDispersionCorrection);
            return s.addForce(nb);
        }, py::arg("charge"), py::arg("sigma"), py::arg("epsilon"), py::arg("exceptionPairs"), py::arg("exceptionParams"),
           py::arg("cutoff"), py::arg("ewaldTolerance") = 5e-4, py::arg("forceGroup") = 0, py::arg("reciprocalSpaceForceGroup") = -1,
           py::arg("useDispersionCorrection") = true)
        .def("addHarmonicBondForce", [](OpenMM::System &s, const std::vector<int> &particle1, const std::vector<int> &particle2,
                                        const std::vector<double> &length, const std::vector<double> &k, int group, bool periodic) {
            if (particle1.size() != particle2.size() || particle1.size() != length.size() || particle1.size() != k.size())
                throw OpenMM::OpenMMException("addHarmonicBondForce: inconsistent array lengths");
            auto *hb = new OpenMM::HarmonicBondForce();
            for (size_t b = 0; b < particle1.size(); b++) {
                if (particle1[b] < 0 || particle2[b] < 0 || particle1[b] >= s.getNumParticles() || particle2[b] >= s.getNumParticles())
                    throw OpenMM::OpenMMException("HarmonicBondForce: Illegal particle index for a bond");
                hb->addBond(particle1[b], particle2[b], length[b], k[b]);
            }
            hb->setForceGroup(group);
            hb->setUsesPeriodicBoundaryConditions(periodic);
            return s.addForce(hb);
        }, py::arg("particle1"), py::arg("particle2"), py::arg("length"), py::arg("k"), py::arg("forceGroup") = 0,
           py::arg("usesPeriodicBoundaryConditions") = false)
        .def("addATMMetaForce", [](OpenMM::System &s, const ATMMetaForce &f) {
            auto *copy = new ATMMetaForce(f);     // the System owns its forces
            s.addForce(copy);
            return copy;
        }, py::return_value_policy::reference_internal, py::arg("force"));

    // ---- platforms: the kernel-less host platform is built in; "CUDA" (the stand-in of OpenMM's CUDA platform) appears
    //      once a plugin library has been loaded / registerATMMetaForceCudaKernelFactories has run
    m.def("loadPluginLibrary", &OpenMM::Platform::loadPluginLibrary, py::arg("file"));
    m.def("registerCudaPlatform", []() {
        try {
            OpenMM::Platform::getPlatformByName("CUDA");
        } catch (const OpenMM::OpenMMException &) {
            OpenMM::Platform::registerPlatform(new OpenMM::CudaPlatform());
        }
    });
    m.def("getPlatformNames", []() {
        std::vector<std::string> names;
        for (int i = 0; i < OpenMM::Platform::getNumPlatforms(); i++) names.push_back(OpenMM::Platform::getPlatform(i).getName());
        return names;
    });
    m.def("platformSupportsKernels", [](const std::string &platform, const std::vector<std::string> &kernels) {
        return OpenMM::Platform::getPlatformByName(platform).supportsKernels(kernels);
    });

    py::class_<OpenMM::Context>(m, "Context")
        .def(py::init<const OpenMM::System &>(), py::keep_alive<1, 2>(), py::arg("system"))
        .def(py::init([](const OpenMM::System &system, const std::string &platform, const std::map<std::string, std::string> &properties) {
            return new OpenMM::Context(system, OpenMM::Platform::getPlatformByName(platform), properties);
        }), py::keep_alive<1, 2>(), py::arg("system"), py::arg("platform"), py::arg("properties") = std::map<std::string, std::string>())
        .def("getPlatformName", [](OpenMM::Context &c) { return c.getPlatform().getName(); })
        .def("usesPlatformKernel", [](OpenMM::Context &c, const ATMMetaForce &f) {
            return dynamic_cast<ATMMetaForceImpl &>(c.getForceImpl(f)).usesPlatformKernel();
        }, py::arg("force"))
        .def("getAtomIndex", [](OpenMM::Context &c) { return OpenMM::CudaPlatform::cudaContext(c.getImpl()).getAtomIndex(); })
        .def("reorderAtoms", [](OpenMM::Context &c, const std::vector<int> &order) {
            OpenMM::CudaPlatform::cudaContext(c.getImpl()).reorderAtoms(order);
        }, py::arg("order"))
        .def("getInnerPosq", [](OpenMM::Context &c, const ATMMetaForce &f, int state) {
            // posq (float4 per slot) of inner context 1 or 2, or of the outer context (state 0): what copyState wrote
            OpenMM::ContextImpl *ci = &c.getImpl();
            if (state != 0) {
                OpenMM::Context *inner = dynamic_cast<ATMMetaForceImpl &>(c.getForceImpl(f)).getInnerContext(state);
                if (!inner) throw OpenMM::OpenMMException("the inner contexts do not exist yet");
                ci = &inner->getImpl();
            }
            OpenMM::CudaContext &cu = OpenMM::CudaPlatform::cudaContext(*ci);
            if (cu.getUseDoublePrecision()) throw OpenMM::OpenMMException("getInnerPosq: single / mixed precision only");
            cu.synchronize();
            py::array_t<float> out({(py::ssize_t)cu.getPaddedNumAtoms(), (py::ssize_t)4});
            cu.getPosq().download(out.mutable_data());
            return out;
        }, py::arg("force"), py::arg("state"))
        .def("setPositions", [](OpenMM::Context &c, py::array_t<double, py::array::c_style | py::array::forcecast> pos) {
            if (pos.ndim() != 2 || pos.shape(1) != 3) throw OpenMM::OpenMMException("setPositions: expected an (N, 3) array in nm");
            std::vector<OpenMM::Vec3> v(pos.shape(0));
            auto r = pos.unchecked<2>();
            for (py::ssize_t i = 0; i < pos.shape(0); i++) v[i] = OpenMM::Vec3(r(i, 0), r(i, 1), r(i, 2));
            c.setPositions(v);
        })
        .def("setPeriodicBoxVectors", [](OpenMM::Context &c, const std::vector<double> &a, const std::vector<double> &b,
                                         const std::vector<double> &cc) { c.setPeriodicBoxVectors(vec3(a), vec3(b), vec3(cc)); })
        .def("setParameter", &OpenMM::Context::setParameter)
        .def("getParameter", &OpenMM::Context::getParameter)
        .def("getParameters", &OpenMM::Context::getParameters)
        .def("setPairListSkins", [](OpenMM::Context &c, const ATMMetaForce &f, double inner, double outer) {
            dynamic_cast<ATMMetaForceImpl &>(c.getForceImpl(f)).setPairListSkins(inner, outer);
        }, py::arg("force"), py::arg("inner"), py::arg("outer"))
        .def("getEnergyRecord", [](OpenMM::Context &c, const ATMMetaForce &f) {
            return dynamic_cast<ATMMetaForceImpl &>(c.getForceImpl(f)).getEnergyRecord();
        }, py::arg("force"))
        .def("calcForcesAndEnergy", [](OpenMM::Context &c, bool includeForces, bool includeEnergy, int groups) {
            const double e = c.calcForcesAndEnergy(includeForces, includeEnergy, groups);
            const auto &f = c.getForces();
            py::array_t<double> out({(py::ssize_t)f.size(), (py::ssize_t)3});
            auto w = out.mutable_unchecked<2>();
            for (size_t i = 0; i < f.size(); i++)
                for (int k = 0; k < 3; k++) w(i, k) = f[i][k];
            return py::make_tuple(e, out);
        }, py::arg("includeForces") = true, py::arg("includeEnergy") = true, py::arg("groups") = -1);

    py::class_<ATMMetaForceB200Kernel>(m, "ATMMetaForceB200Kernel")
        .def(py::init<>())
        .def_static("Name", &ATMMetaForceB200Kernel::Name)
        .def("initialize", [](ATMMetaForceB200Kernel &k, const ATMMetaForce &f, int padded, int precision,
                              const std::vector<int> &atomIndex, int device) {
            k.initialize(f, padded, (atm_precision)precision, atomIndex, device);
        }, py::arg("force"), py::arg("paddedNumAtoms"), py::arg("precision"), py::arg("atomIndex") = std::vector<int>(),
           py::arg("device") = -1)
        .def("atomsReordered", [](ATMMetaForceB200Kernel &k, const ATMMetaForce &f, const std::vector<int> &idx, uintptr_t stream) {
            k.atomsReordered(f, idx, ptr(stream));
        }, py::arg("force"), py::arg("atomIndex"), py::arg("stream") = 0)
        .def("copyState", [](ATMMetaForceB200Kernel &k, uintptr_t posq, uintptr_t corr, uintptr_t posq1, uintptr_t corr1,
                             uintptr_t posq2, uintptr_t corr2, uintptr_t stream) {
            k.copyState(ptr(posq), ptr(corr), ptr(posq1), ptr(corr1), ptr(posq2), ptr(corr2), ptr(stream));
        }, py::arg("posq"), py::arg("posqCorrection"), py::arg("posq1"), py::arg("posq1Correction"), py::arg("posq2"),
           py::arg("posq2Correction"), py::arg("stream") = 0)
        .def("execute", [](ATMMetaForceB200Kernel &k, const std::map<std::string, double> &params, uintptr_t force,
                           uintptr_t f1, uintptr_t f2, double U1, double U2, bool includeForces, bool includeEnergy,
                           uintptr_t stream) {
            return k.execute(params, (long long *)ptr(force), (const long long *)ptr(f1), (const long long *)ptr(f2), U1, U2,
                             includeForces, includeEnergy, ptr(stream));
        }, py::arg("parameters"), py::arg("force"), py::arg("forceState1"), py::arg("forceState2"), py::arg("State1Energy"),
           py::arg("State2Energy"), py::arg("includeForces") = true, py::arg("includeEnergy") = true, py::arg("stream") = 0)
        .def("copyParametersToContext", [](ATMMetaForceB200Kernel &k, const ATMMetaForce &f, uintptr_t stream) {
            k.copyParametersToContext(f, ptr(stream));
        }, py::arg("force"), py::arg("stream") = 0)
        .def("getPerturbationEnergy", &ATMMetaForceB200Kernel::getPerturbationEnergy)
        .def_static("getDefaultParameters", &ATMMetaForceB200Kernel::getDefaultParameters);
}
