"""ATMMetaForce -- Python face of the C++ class (pybind11 binding of openmmapi/include/ATMMetaForce.h), with the
conveniences the reference's SWIG layer adds (ref: python/atmmetaforceplugin.i): Quantity arguments are stripped to
the MD unit system (nm, kJ/mol, ps), getPerturbationEnergy(context) is unit-decorated when openmm.unit is available,
and cast()/isinstance() helpers exist."""
from . import _atmmetaforce_core as _core

try:  # unit decoration only when OpenMM's unit package is present (it is absent in the build image)
    from openmm import unit as _unit
except Exception:  # pragma: no cover
    _unit = None

OpenMMException = _core.OpenMMException


def _strip(x):
    """Quantity -> float in the MD unit system; plain numbers pass through (duck-typed like OpenMM's typemaps)."""
    if hasattr(x, "value_in_unit_system"):
        from openmm.unit import md_unit_system
        return x.value_in_unit_system(md_unit_system)
    return float(x)


class vectord(list):
    """std::vector<double> as SWIG exposes it (ref: python/atmmetaforceplugin.i:16 %template(vectord)): a sequence of
    floats with push_back / size besides the list protocol."""

    _conv = float

    def __init__(self, values=()):
        if isinstance(values, int) and not isinstance(values, bool):   # vectord(n): n default-constructed elements
            values = [self._conv(0)] * values
        super().__init__(self._conv(_strip(v) if self._conv is float else v) for v in values)

    def append(self, v):
        super().append(self._conv(_strip(v) if self._conv is float else v))

    push_back = append

    def __setitem__(self, i, v):
        if isinstance(i, slice):
            super().__setitem__(i, [self._conv(x) for x in v])
        else:
            super().__setitem__(i, self._conv(_strip(v) if self._conv is float else v))

    def size(self):
        return len(self)

    def empty(self):
        return len(self) == 0


class vectori(vectord):
    """std::vector<int> (ref: .i:17 %template(vectori)); what getVariableForceGroups() hands back."""

    _conv = int


class ATMMetaForce(_core.ATMMetaForce):
    """See openmmapi/include/ATMMetaForce.h.  Constructor order (ref: ATMMetaForce.h:79-83):
    (Lambda1, Lambda2, Alpha, U0, W0, Umax, Ubcore, Acore, direction, VariableForceGroups)."""

    def __init__(self, Lambda1, Lambda2, Alpha, U0, W0, Umax, Ubcore, Acore, direction, VariableForceGroups):
        super().__init__(_strip(Lambda1), _strip(Lambda2), _strip(Alpha), _strip(U0), _strip(W0), _strip(Umax),
                         _strip(Ubcore), _strip(Acore), _strip(direction), [int(g) for g in VariableForceGroups])

    def addParticle(self, particle, dx, dy, dz):
        return super().addParticle(int(particle), _strip(dx), _strip(dy), _strip(dz))

    def setParticleParameters(self, index, particle, dx, dy, dz):
        return super().setParticleParameters(int(index), int(particle), _strip(dx), _strip(dy), _strip(dz))

    def getPerturbationEnergy(self, context):
        """Soft-core perturbation energy of the context's last evaluation, kJ/mol
        (ref: ATMMetaForce::getPerturbationEnergy, openmmapi/src/ATMMetaForce.cpp:42-44; .i:47-49 adds the unit)."""
        if isinstance(context, _core.Context):      # the C++ Context of this build: ATMMetaForceImpl answers
            val = super().getPerturbationEnergy(context)
        else:
            val = context._atm_perturbation_energy(self)
        return val * _unit.kilojoules_per_mole if _unit is not None else val

    def updateParametersInContext(self, context):
        """Push changed displacements to the device (ref: ATMMetaForce.cpp:38-40)."""
        if isinstance(context, _core.Context):
            super().updateParametersInContext(context)
        else:
            context._atm_update_parameters(self)

    def getVariableForceGroups(self):
        return vectori(super().getVariableForceGroups())

    @staticmethod
    def cast(force):
        if not isinstance(force, _core.ATMMetaForce):
            raise TypeError("not an ATMMetaForce")
        return force

    @staticmethod
    def isinstance(force):
        return isinstance(force, _core.ATMMetaForce)


def serialize(force, rootName="Force"):
    """XmlSerializer.serialize for an ATMMetaForce (reference schema, serialization/src/ATMMetaForceProxy.cpp)."""
    return _core.serialize(force, rootName)


def deserialize(xml):
    c = _core.deserialize(xml)  # the C++ proxy validates the document and rebuilds the C++ object
    f = ATMMetaForce(*c.getDefaultParameterArray(), c.getVariableForceGroups())
    f.setForceGroup(c.getForceGroup())
    f.setName(c.getName())
    for i in range(c.getNumParticles()):
        f.addParticle(*c.getParticleParameters(i))
    return f
