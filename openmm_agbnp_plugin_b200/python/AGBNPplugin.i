/* SWIG interface of the Python module "AGBNPplugin": the public surface of the plugin (module name, class name, method
 * names, enum and argument order) is the reference's (python/AGBNPPlugin.i:1,47-85), so that scripts written for it --
 * example/test_agbnp.py, example/*_benchmark.py -- run unchanged on top of the B200 back end:
 *
 *     from AGBNPplugin import AGBNPForce
 *     gb = AGBNPForce(); gb.setNonbondedMethod(CutoffNonPeriodic); gb.setCutoffDistance(1.2 * nanometer); gb.setVersion(1)
 *     gb.addParticle(radius, gamma, alpha, charge, ishydrogen)
 *
 * Without OpenMM (this repository's tests) the pure-Python mirror openmm_agbnp_plugin_b200/AGBNPplugin.py offers the same
 * class over the C-ABI.
 */
%module AGBNPplugin

/* AGBNPForce derives from OpenMM::Force: take Force, Context, ... from the OpenMM module (openmm >= 7.6 installs "openmm",
 * older releases "simtk.openmm") */
%import(module="openmm") "swig/OpenMMSwigHeaders.i"
%include "swig/typemaps.i"
%include "std_vector.i"
namespace std {
    %template(vectord) vector<double>;
    %template(vectori) vector<int>;
};

%{
#include "AGBNPForce.h"
#include "OpenMM.h"
#include "OpenMMAmoeba.h"
#include "OpenMMDrude.h"
#include "openmm/RPMDIntegrator.h"
#include "openmm/RPMDMonteCarloBarostat.h"
%}

%pythoncode %{
try:
    import openmm as mm
    import openmm.unit as unit
except ImportError:                     # OpenMM < 7.6
    import simtk.openmm as mm
    import simtk.unit as unit
%}

/* getParticleParameters returns (radius, gamma, alpha, charge, ishydrogen); radius and gamma carry units */
%pythonappend AGBNPPlugin::AGBNPForce::getParticleParameters(int index, double& radius, double& gamma, double& alpha,
                                                             double& charge, bool& ishydrogen) const %{
    val = list(val)
    val[0] = unit.Quantity(val[0], unit.nanometer)
    val[1] = unit.Quantity(val[1], unit.kilojoule_per_mole / (unit.nanometer * unit.nanometer))
%}

namespace AGBNPPlugin {

class AGBNPForce : public OpenMM::Force {
public:
    AGBNPForce();

    int getNumParticles() const;
    int addParticle(double radius, double gamma, double alpha, double charge, bool ishydrogen);
    void setParticleParameters(int index, double radius, double gamma, double alpha, double charge, bool ishydrogen);
    void updateParametersInContext(OpenMM::Context& context);

    enum NonbondedMethod {
        NoCutoff = 0,
        CutoffNonPeriodic = 1,
        CutoffPeriodic = 2
    };
    NonbondedMethod getNonbondedMethod() const;
    void setNonbondedMethod(NonbondedMethod method);
    double getCutoffDistance() const;
    void setCutoffDistance(double distance);

    unsigned int getVersion() const;
    void setVersion(int agbnp_version);

    /* reference parameters are outputs: SWIG returns them as a tuple */
    %apply double& OUTPUT {double& radius};
    %apply double& OUTPUT {double& gamma};
    %apply double& OUTPUT {double& alpha};
    %apply double& OUTPUT {double& charge};
    %apply bool& OUTPUT {bool& ishydrogen};
    void getParticleParameters(int index, double& radius, double& gamma, double& alpha, double& charge, bool& ishydrogen) const;
    %clear double& radius;
    %clear double& gamma;
    %clear double& alpha;
    %clear double& charge;
    %clear bool& ishydrogen;
};

}
