"""A minimal stand-alone Context for the ATM path (OpenMM itself cannot be installed in this image).

It plays the role OpenMM's Context + ATMMetaForceImpl play around the reference plugin
(ref: openmmapi/src/ATMMetaForceImpl.cpp:69-142): it owns the nine global parameters under their ATM* names,
positions, box, and runs  copy-state -> two-state direct space -> scalar stage -> merge  on the GPU through the
C ABI.  The variable force group is a NonbondedForce-like description (charges, sigma, epsilon, exclusions,
exceptions, cutoff, Ewald alpha) of the System.  As in the reference's inner contexts the NonbondedForce is evaluated as a
whole by default: direct space, PME reciprocal space of both states (OpenMM's mesh rule, order 5) and the long-range
dispersion correction; `reciprocal_space=False` leaves the reciprocal part to an external evaluator
(setExternalStateEnergies), which is what north_star's Tier-2 integration does with OpenMM's own PME.
"""
import numpy as np

from . import _capi
from .backend import ATMBackend
from .force import ATMMetaForce


class NonbondedDirect:
    """The part of OpenMM's NonbondedForce this back-end evaluates (what setNonbondedMethod(PME) computes in direct
    space).  Arrays by atom; exceptions are (i, j, chargeProd, sigma, epsilon) rows and imply an exclusion."""

    def __init__(self, charge, sigma, epsilon, cutoff, ewald_alpha=None, ewald_tolerance=5e-4, exclusions=None,
                 exception_pairs=None, exception_params=None, force_group=0, reciprocal_space=True, dispersion_correction=True):
        self.charge = np.ascontiguousarray(charge, np.float64)
        self.sigma = np.ascontiguousarray(sigma, np.float64)
        self.epsilon = np.ascontiguousarray(epsilon, np.float64)
        self.cutoff = float(cutoff)
        self.ewald_alpha = float(ewald_alpha) if ewald_alpha is not None else float(np.sqrt(-np.log(2 * ewald_tolerance)) / cutoff)
        self.exclusions = np.zeros((0, 2), np.int32) if exclusions is None else np.ascontiguousarray(exclusions, np.int32)
        self.exception_pairs = np.zeros((0, 2), np.int32) if exception_pairs is None else np.ascontiguousarray(exception_pairs, np.int32)
        self.exception_params = np.zeros((0, 3)) if exception_params is None else np.ascontiguousarray(exception_params, np.float64)
        self.force_group = int(force_group)
        self.ewald_tolerance = float(ewald_tolerance)
        self.reciprocal_space = bool(reciprocal_space)            # NonbondedForce.setReciprocalSpaceForceGroup(-1): part of the force
        self.dispersion_correction = bool(dispersion_correction)  # NonbondedForce.setUseDispersionCorrection (OpenMM default: on)

    def getForceGroup(self):
        return self.force_group

    def setForceGroup(self, g):
        self.force_group = int(g)


class State:
    def __init__(self, energy, forces, positions):
        self._e, self._f, self._p = energy, forces, positions

    def getPotentialEnergy(self):
        return self._e

    def getForces(self, asNumpy=True):
        return self._f

    def getPositions(self, asNumpy=True):
        return self._p


class Context:
    """Context(atmforce, nonbonded, box) -- evaluates the ATM force group on cuda:`device`."""

    PARAM_ORDER = ("ATMLambda1", "ATMLambda2", "ATMAlpha", "ATMU0", "ATMW0", "ATMUmax", "ATMUbcore", "ATMAcore", "ATMDirection")

    def __init__(self, atmforce, nonbonded, box, precision="mixed", skin=0.1, skin_outer=0.3, device=0):
        import torch
        if not isinstance(atmforce, ATMMetaForce):
            raise TypeError("atmforce must be an ATMMetaForce")
        if nonbonded.getForceGroup() not in atmforce.getVariableForceGroups():
            raise _capi.ATMError("the NonbondedForce group is not one of the ATM variable force groups: nothing to evaluate")
        if atmforce.getForceGroup() in atmforce.getVariableForceGroups():
            # ref: ATMMetaForceImpl.cpp:78-79
            raise _capi.ATMError("The ATM Meta Force group cannot be one of the variable force groups.")
        self._torch = torch
        self._force = atmforce
        self._nb = nonbonded
        n = atmforce.getNumParticles()
        if n != nonbonded.charge.size:
            raise _capi.ATMError("ATMMetaForce and the NonbondedForce must have one entry per System particle")
        self._n = n
        self._be = ATMBackend(n, precision=precision, num_replicas=1, device=device)
        self._P = self._be.P
        self._params = dict(zip(self.PARAM_ORDER, atmforce.getDefaultParameterArray()))  # ref: ATMMetaForceImpl.cpp:130-142
        self._box = np.asarray(box, np.float64)
        self._be.set_displacements(np.asarray(atmforce.getDisplacementArray()).reshape(n, 3))
        self._be.set_box(self._box)
        self._be.nb_setup(nonbonded.charge, nonbonded.sigma, nonbonded.epsilon, nonbonded.cutoff, nonbonded.ewald_alpha,
                          skin=skin, skin_outer=skin_outer, exclusions=nonbonded.exclusions,
                          exception_pairs=nonbonded.exception_pairs, exception_params=nonbonded.exception_params)
        self._skin, self._skin_outer = skin, skin_outer
        self._setup_reciprocal()
        self._dev = torch.device("cuda", device)
        self._posq = torch.zeros((1, self._P, 4), dtype=torch.float32, device=self._dev)
        self._corr = torch.zeros((1, self._P, 4), dtype=torch.float32, device=self._dev)
        self._pos64 = None
        self._ref_rebuild = None
        self._ref_prune = None
        self._energy_ext = None
        self._last = None

    def _setup_reciprocal(self):
        from .synthetic import pme_grid
        if self._nb.reciprocal_space:
            self._be.pme_setup(pme_grid(np.diag(self._box) if self._box.ndim == 2 else self._box, self._nb.ewald_alpha,
                                        self._nb.ewald_tolerance), 5)
        self._be.set_dispersion_correction(self._nb.dispersion_correction)

    # -- OpenMM-like surface ------------------------------------------------------------------------------------
    def setPositions(self, positions):
        pos = np.asarray(positions.value_in_unit_system(__import__("openmm").unit.md_unit_system)
                         if hasattr(positions, "value_in_unit_system") else positions, np.float64).reshape(self._n, 3)
        p32 = pos.astype(np.float32)
        posq = np.zeros((1, self._P, 4), np.float32)
        posq[0, :self._n, :3] = p32
        posq[0, :self._n, 3] = self._nb.charge
        corr = np.zeros_like(posq)
        corr[0, :self._n, :3] = (pos - p32).astype(np.float32)
        self._posq.copy_(self._torch.from_numpy(posq))
        self._corr.copy_(self._torch.from_numpy(corr))
        self._pos64 = pos

    def setPeriodicBoxVectors(self, a, b=None, c=None):
        box = np.asarray(a, np.float64) if b is None else np.array([a, b, c], np.float64)
        self._box = box
        self._be.set_box(box)
        self._setup_reciprocal()
        self._ref_rebuild = None

    def setParameter(self, name, value):
        if name not in self._params:
            raise _capi.ATMError(f"Called setParameter() with invalid parameter name: {name}")
        self._params[name] = float(value.value_in_unit_system(__import__("openmm").unit.md_unit_system)
                                   if hasattr(value, "value_in_unit_system") else value)

    def getParameter(self, name):
        if name not in self._params:
            raise _capi.ATMError(f"Called getParameter() with invalid parameter name: {name}")
        return self._params[name]

    def getParameters(self):
        return dict(self._params)

    def setExternalStateEnergies(self, U1_ext, U2_ext):
        """Energies of the variable force groups evaluated elsewhere (e.g. OpenMM's PME reciprocal space in the inner
        contexts); added to U1/U2 on the device before the scalar stage."""
        self._energy_ext = self._torch.tensor([[float(U1_ext), float(U2_ext)]], dtype=self._torch.float64, device=self._dev)

    def getState(self, getEnergy=False, getForces=False, getPositions=False, groups=-1):
        """Evaluates the ATM force group (ref: ATMMetaForceImpl::calcForcesAndEnergy, ATMMetaForceImpl.cpp:90-128)."""
        if self._pos64 is None:
            raise _capi.ATMError("Particle positions have not been set")
        atm_group = self._force.getForceGroup()
        in_groups = groups == -1 or (isinstance(groups, (set, list, tuple)) and atm_group in groups) or \
            (isinstance(groups, int) and groups != -1 and (groups >> atm_group) & 1)
        energy, forces = 0.0, np.zeros((self._n, 3))
        if in_groups:
            self._maintain_lists()
            self._be.set_parameters([self._params[k] for k in self.PARAM_ORDER])
            f = self._torch.zeros((1, 3 * self._P), dtype=self._torch.int64, device=self._dev)
            self._be.step(self._posq, f, posq_corr=self._corr, energy_ext=self._energy_ext, include_energy=True)
            en = self._be.get_energies()[0]
            self._last = en
            energy = float(en[_capi.E_ENERGY]) if getEnergy else 0.0
            if getForces:
                forces = f.cpu().numpy()[0].reshape(3, self._P)[:, :self._n].T.astype(np.float64) / 4294967296.0
        return State(energy, forces if getForces else None, self._pos64 if getPositions else None)

    # -- hooks used by ATMMetaForce --------------------------------------------------------------------------------
    def _atm_perturbation_energy(self, force):
        if self._last is None:
            return 0.0
        return float(self._last[_capi.E_USC])

    def _atm_update_parameters(self, force):
        if force.getNumParticles() != self._n:
            raise _capi.ATMError("copyParametersToContext: The number of ATMMetaForce particles has changed")
        self._be.set_displacements(np.asarray(force.getDisplacementArray()).reshape(self._n, 3))
        self._ref_rebuild = None

    def _maintain_lists(self):
        """Rebuild / prune the pair lists when atoms moved by more than half the respective skin."""
        pos = self._pos64
        if self._ref_rebuild is None or np.abs(pos - self._ref_rebuild).max() * np.sqrt(3.0) > 0.5 * self._skin_outer:
            self._be.rebuild(self._posq)
            self._ref_rebuild = pos.copy()
            self._ref_prune = pos.copy()
        elif np.abs(pos - self._ref_prune).max() * np.sqrt(3.0) > 0.5 * self._skin:
            self._be.prune(self._posq)
            self._ref_prune = pos.copy()

    def close(self):
        self._be.close()
