"""On-disk formats either side of the hot path.

* `.out` sample lines -- one line per sample, `T lambda lambda1 lambda2 alpha u0 w0 PE u` in kcal/mol units, exactly the
  line the reference's example scripts print (ref: example/abfe/abfe.py:149-160, README.md:193-201).
* OpenMM State XML -- the checkpoint format the examples load and save (ref: example/abfe/temoa-g1-equil.xml:1-10):
  `<State openmmVersion time type="State" version="1">` with `<PeriodicBoxVectors>`, `<Parameters .../>` (one attribute
  per global parameter, the nine ATM* ones included), `<Positions>`, `<Velocities>`, `<IntegratorParameters/>`.
"""
import re
import xml.etree.ElementTree as ET

import numpy as np

KCAL = 4.184  # kJ per kcal


def format_sample_line(temperature, lmbd, lambda1, lambda2, alpha, u0, w0, pot_energy, pert_energy):
    """All energies in kJ/mol (alpha in (kJ/mol)^-1) on input; the line is written in kcal/mol units like the reference."""
    return "%f %f %f %f %f %f %f %f %f" % (temperature, lmbd, lambda1, lambda2, alpha * KCAL, u0 / KCAL, w0 / KCAL,
                                           pot_energy / KCAL, pert_energy / KCAL)


def parse_sample_line(line):
    """Inverse of format_sample_line (values back in kJ/mol)."""
    f = [float(x) for x in line.split()]
    if len(f) != 9:
        raise ValueError("a sample line has 9 columns: T lambda lambda1 lambda2 alpha u0 w0 PE u")
    return dict(temperature=f[0], lmbd=f[1], lambda1=f[2], lambda2=f[3], alpha=f[4] / KCAL, u0=f[5] * KCAL, w0=f[6] * KCAL,
                pot_energy=f[7] * KCAL, pert_energy=f[8] * KCAL)


class SampleWriter:
    """`<job>.out` writer: one flushed line per sample."""

    def __init__(self, path, temperature=300.0):
        self._f = open(path, "w")
        self.temperature = float(temperature)

    def write(self, context, atmforce, pot_energy, lmbd=None):
        l1, l2 = context.getParameter(atmforce.Lambda1()), context.getParameter(atmforce.Lambda2())
        pert = atmforce.getPerturbationEnergy(context)
        pert = pert.value_in_unit(pert.unit) if hasattr(pert, "value_in_unit") else pert
        line = format_sample_line(self.temperature, l2 if lmbd is None else lmbd, l1, l2, context.getParameter(atmforce.Alpha()),
                                  context.getParameter(atmforce.U0()), context.getParameter(atmforce.W0()), pot_energy, pert)
        self._f.write(line + "\n")
        self._f.flush()
        return line

    def close(self):
        self._f.close()


def _fmt(x):
    """OpenMM prints the shortest round-trip decimal and drops a leading zero ('.0625')."""
    s = repr(float(x))
    if "e" in s or "inf" in s or "nan" in s:
        return s
    if s.endswith(".0"):
        s = s[:-2]
    if s.startswith("0."):
        s = s[1:]
    elif s.startswith("-0."):
        s = "-" + s[2:]
    return s


def write_state_xml(path, positions, box, parameters, velocities=None, time=0.0, openmm_version="7.7"):
    """positions/velocities (N,3) in nm, nm/ps; box (3,) or (3,3) in nm; parameters: dict name -> value."""
    box = np.asarray(box, np.float64)
    if box.size == 3:
        box = np.diag(box)
    out = ['<?xml version="1.0" ?>', f'<State openmmVersion="{openmm_version}" time="{_fmt(time)}" type="State" version="1">',
           "\t<PeriodicBoxVectors>"]
    for tag, v in zip("ABC", box):
        out.append(f'\t\t<{tag} x="{_fmt(v[0])}" y="{_fmt(v[1])}" z="{_fmt(v[2])}"/>')
    out.append("\t</PeriodicBoxVectors>")
    out.append("\t<Parameters " + " ".join(f'{k}="{_fmt(parameters[k])}"' for k in sorted(parameters)) + "/>")
    out.append("\t<Positions>")
    for p in np.asarray(positions, np.float64):
        out.append(f'\t\t<Position x="{_fmt(p[0])}" y="{_fmt(p[1])}" z="{_fmt(p[2])}"/>')
    out.append("\t</Positions>")
    if velocities is not None:
        out.append("\t<Velocities>")
        for v in np.asarray(velocities, np.float64):
            out.append(f'\t\t<Velocity x="{_fmt(v[0])}" y="{_fmt(v[1])}" z="{_fmt(v[2])}"/>')
        out.append("\t</Velocities>")
    out.append("\t<IntegratorParameters/>")
    out.append("</State>")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


def read_state_xml(path):
    """Returns dict(time, box (3,3), parameters {name: float}, positions (N,3), velocities (N,3) or None)."""
    root = ET.parse(path).getroot()
    if root.tag != "State" or root.attrib.get("type", "State") != "State":
        raise ValueError("not an OpenMM State XML file")

    def vec(e):
        return [float(e.attrib["x"]), float(e.attrib["y"]), float(e.attrib["z"])]
    pbv = root.find("PeriodicBoxVectors")
    box = np.array([vec(pbv.find(t)) for t in "ABC"]) if pbv is not None else None
    par = root.find("Parameters")
    params = {k: float(v) for k, v in par.attrib.items()} if par is not None else {}
    pos = root.find("Positions")
    positions = np.array([vec(e) for e in pos]) if pos is not None else None
    vel = root.find("Velocities")
    velocities = np.array([vec(e) for e in vel]) if vel is not None and len(vel) else None
    return dict(time=float(root.attrib.get("time", 0.0)), box=box, parameters=params, positions=positions,
                velocities=velocities, openmm_version=root.attrib.get("openmmVersion"))


def load_state(context, path):
    """simulation.loadState for the stand-alone Context: positions, box and every global parameter the context knows
    (files written before a parameter existed -- e.g. ATMDirection in the reference's fixtures -- leave it at its default)."""
    st = read_state_xml(path)
    if st["box"] is not None:
        context.setPeriodicBoxVectors(st["box"])
    context.setPositions(st["positions"])
    known = context.getParameters()
    for k, v in st["parameters"].items():
        if k in known:
            context.setParameter(k, v)
    return st
