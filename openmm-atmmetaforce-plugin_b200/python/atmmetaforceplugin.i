/* atmmetaforceplugin.i -- SWIG interface of the Python module `atmmetaforce` for builds WITH OpenMM and SWIG.
 *
 * Same module name, class surface, std::vector template names (vectord, vectori), OUTPUT tuple of
 * getParticleParameters, unit decoration and exception mapping as the reference's interface file
 * (ref: python/atmmetaforceplugin.i:1-126), written against this repository's openmmapi/include/ATMMetaForce.h.
 * SWIG and OpenMM's swig headers are absent from the build image of this repository, so this file is not generated
 * here; the same surface is provided without SWIG by the pybind11 module python/src/atmmetaforce_core.cpp together with
 * python/atmmetaforce/force.py (tests/test_facade.py exercises that surface, including vectord / vectori).
 *
 *   swig -python -c++ -I$OPENMM_DIR/include -Iopenmm-atmmetaforce-plugin_b200/openmmapi/include \
 *        -o ATMMetaForcePluginWrapper.cpp python/atmmetaforceplugin.i
 *   (compile the wrapper with -DATM_HAVE_OPENMM and link -lATMMetaForcePlugin -lOpenMM)
 */
%module atmmetaforce

%include "factory.i"
%import(module="openmm") "swig/OpenMMSwigHeaders.i"
%include "swig/typemaps.i"
%include <std_string.i>
%include <std_vector.i>

/* the two vector types the constructor and getVariableForceGroups() exchange with Python */
namespace std {
  %template(vectord) vector<double>;
  %template(vectori) vector<int>;
};

%{
#define ATM_HAVE_OPENMM
#include "ATMMetaForce.h"
#include "OpenMM.h"
%}

%pythoncode %{
import math
import openmm as mm
from openmm.unit import *
%}

%include "ATMMetaForceVersion.h"

/* ATMMetaForceUtils travels inside the generated module, as in the reference (utils.py of this repository is written
 * for exactly that: it only needs `mm`, the unit names and `math` in the module namespace) */
%pythoncode "atmmetaforce/utils.py"

/* displacements come back in nanometres, the perturbation energy in kJ/mol */
%pythonappend ATMMetaForcePlugin::ATMMetaForce::getParticleParameters(int index, int& particle, double& dx, double& dy, double& dz) const %{
    val = [val[0], Quantity(val[1], nanometer), Quantity(val[2], nanometer), Quantity(val[3], nanometer)]
%}
%pythonappend ATMMetaForcePlugin::ATMMetaForce::getPerturbationEnergy(OpenMM::Context& context) const %{
    val = Quantity(val, kilojoules_per_mole)
%}

/* any std::exception (OpenMMException included) becomes a Python Exception carrying what() */
%exception {
    try {
        $action
    } catch (const std::exception& e) {
        PyErr_SetString(PyExc_Exception, e.what());
        SWIG_fail;
    }
}

namespace ATMMetaForcePlugin {

class ATMMetaForce : public OpenMM::Force {
public:
    ATMMetaForce(double Lambda1, double Lambda2, double Alpha, double U0, double W0, double Umax, double Ubcore, double Acore,
                 double direction, const std::vector<int>& VariableForceGroups);

    int getNumParticles() const;
    int addParticle(int particle, double dx, double dy, double dz);
    void setParticleParameters(int index, int particle, double dx, double dy, double dz);

    /* the four reference arguments are results: SWIG returns them as a tuple (particle, dx, dy, dz) */
    %apply int& OUTPUT { int& particle };
    %apply double& OUTPUT { double& dx, double& dy, double& dz };
    void getParticleParameters(int index, int& particle, double& dx, double& dy, double& dz) const;
    %clear int& particle;
    %clear double& dx, double& dy, double& dz;

    void updateParametersInContext(OpenMM::Context& context);
    double getPerturbationEnergy(OpenMM::Context& context) const;
    bool usesPeriodicBoundaryConditions() const;

    /* names of the nine global Context parameters, and the plugin version */
    static const std::string& Lambda1();
    static const std::string& Lambda2();
    static const std::string& Alpha();
    static const std::string& U0();
    static const std::string& W0();
    static const std::string& Umax();
    static const std::string& Ubcore();
    static const std::string& Acore();
    static const std::string& Direction();
    static const std::string& Version();

    double getDefaultLambda1() const;
    double getDefaultLambda2() const;
    double getDefaultAlpha() const;
    double getDefaultU0() const;
    double getDefaultW0() const;
    double getDefaultUmax() const;
    double getDefaultUbcore() const;
    double getDefaultAcore() const;
    double getDefaultDirection() const;
    const std::vector<int>& getVariableForceGroups() const;

    /* down-casts for Forces that come back from System.getForce() typed as OpenMM::Force */
    %extend {
        static ATMMetaForcePlugin::ATMMetaForce& cast(OpenMM::Force& force) {
            return dynamic_cast<ATMMetaForcePlugin::ATMMetaForce&>(force);
        }
        static bool isinstance(OpenMM::Force& force) {
            return dynamic_cast<ATMMetaForcePlugin::ATMMetaForce*>(&force) != NULL;
        }
    }
};

}  /* namespace ATMMetaForcePlugin */
