// pybind11 binding of the C++ facade (plays the role of the reference's SWIG wrapper, python/atmmetaforceplugin.i,
// which cannot be generated here: SWIG and OpenMM's swig headers are absent).  Same method names and argument order;
// getParticleParameters returns the tuple (particle, dx, dy, dz) like the SWIG OUTPUT typemaps (.i:79-87);
// std::exception is mapped to a Python exception like the %exception block (.i:54-61).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>

#include "ATMMetaForce.h"
#include "ATMMetaForceB200Kernel.h"
#include "ATMMetaForceProxy.h"

namespace py = pybind11;
using namespace ATMMetaForcePlugin;

static void *ptr(uintptr_t p) { return reinterpret_cast<void *>(p); }

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
