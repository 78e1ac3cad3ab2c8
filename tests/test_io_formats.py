"""CPU tests of the on-disk formats either side of the path: `.out` sample lines (ref: example/abfe/abfe.py:149-160,
README.md:193-201) and the OpenMM State XML checkpoint (ref: example/abfe/temoa-g1-equil.xml:1-10)."""
import numpy as np

import atmmetaforce as atm
from atmmetaforce import io


def test_sample_line_matches_reference_layout():
    # README.md:195: "300.000000 0.500000 0.500000 0.500000 0.000000 0.000000 0.000000 -69967.037957 -0.956023"
    line = io.format_sample_line(300.0, 0.5, 0.5, 0.5, 0.0, 0.0, 0.0, -69967.037957 * 4.184, -0.956023 * 4.184)
    assert line == "300.000000 0.500000 0.500000 0.500000 0.000000 0.000000 0.000000 -69967.037957 -0.956023"
    back = io.parse_sample_line(line)
    assert abs(back["pot_energy"] + 69967.037957 * 4.184) < 1e-6 and back["lambda1"] == 0.5
    # alpha is written in (kcal/mol)^-1, u0 / w0 in kcal/mol
    line = io.format_sample_line(300.0, 0.2, 0.0, 0.2, 0.1 / 4.184, 110 * 4.184, 2 * 4.184, 0.0, 0.0)
    assert line.split()[4:7] == ["0.100000", "110.000000", "2.000000"]


def test_state_xml_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    pos = rng.uniform(-5, 9, (50, 3))
    vel = rng.normal(0, 0.5, (50, 3))
    params = {"ATMAcore": 0.0625, "ATMAlpha": 0.0, "ATMLambda1": 0.5000000000000003, "ATMLambda2": 0.5000000000000003,
              "ATMU0": 0.0, "ATMUbcore": 2092.0, "ATMUmax": 4184.0, "ATMW0": 0.0, "MonteCarloPressure": 1.0}
    p = tmp_path / "state.xml"
    io.write_state_xml(p, pos, [4.217498273277886, 4.659005774446288, 4.3635687055440755], params, velocities=vel,
                       time=399.99999999745313, openmm_version="7.6")
    text = p.read_text()
    assert text.splitlines()[1] == '<State openmmVersion="7.6" time="399.99999999745313" type="State" version="1">'
    assert 'ATMAcore=".0625"' in text and 'ATMUbcore="2092"' in text        # OpenMM's number style
    st = io.read_state_xml(p)
    assert np.array_equal(st["positions"], pos) and np.array_equal(st["velocities"], vel)
    assert st["parameters"] == params and st["time"] == 399.99999999745313
    assert np.array_equal(np.diag(st["box"]), [4.217498273277886, 4.659005774446288, 4.3635687055440755])


def test_reads_reference_style_header(tmp_path):
    p = tmp_path / "ref.xml"
    p.write_text('<?xml version="1.0" ?>\n<State openmmVersion="7.6" time="150.00000000035217" type="State" version="1">\n'
                 '\t<PeriodicBoxVectors>\n\t\t<A x="5.748062944833323" y="0" z="0"/>\n\t\t<B x="0" y="6.053858258816803" z="0"/>\n'
                 '\t\t<C x="0" y="0" z="6.254613619984414"/>\n\t</PeriodicBoxVectors>\n'
                 '\t<Parameters ATMAcore=".0625" ATMAlpha="0" ATMLambda1=".5" ATMLambda2=".5" ATMU0="0" '
                 'ATMUbcore="209.20000000000002" ATMUmax="418.40000000000003" ATMW0="0" MonteCarloPressure="1" MonteCarloTemperature="300"/>\n'
                 '\t<Positions>\n\t\t<Position x="1.9480619430541992" y="2.921112298965454" z="2.4037277698516846"/>\n\t</Positions>\n'
                 '\t<Velocities>\n\t\t<Velocity x=".1" y="-.2" z=".3"/>\n\t</Velocities>\n\t<IntegratorParameters/>\n</State>\n')
    st = io.read_state_xml(p)
    assert st["parameters"]["ATMUbcore"] == 209.20000000000002 and "ATMDirection" not in st["parameters"]
    assert st["positions"].shape == (1, 3) and st["velocities"][0, 1] == -0.2 and st["box"][2, 2] == 6.254613619984414
