"""The driver's input / output formats (reference src/OpenMps/Main.cpp:31-274) through the real
executable: `OpenMps --check-io` parses the XML run description and writes the INPUT state as a result CSV without a GPU.
Known answers: the C `%g` text the reference's `ostream << double` produces, and the reference's own error messages."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmps_b200 import scenes  # noqa: E402

BIN2 = os.path.join(ROOT, "openmps_b200", "bin", "OpenMps")
BIN3 = os.path.join(ROOT, "openmps_b200", "bin", "OpenMps3d")

pytestmark = pytest.mark.skipif(not os.path.exists(BIN2), reason="driver not built (python -c 'import __graft_entry__ as g; g.build()')")


def run(binary, *args, cwd=None):
    return subprocess.run([binary, *args], capture_output=True, text=True, cwd=cwd, timeout=120)


def expected_csv(sc):
    head = "Type, x, z, u, w, p, n" if sc.env.dim == 2 else "Type, x, y, z, u, v, w, p, n"
    lines = [head]
    for i in range(sc.count):
        vals = [str(int(sc.type[i]))] + ["%g" % v for v in sc.x[i]] + ["%g" % v for v in sc.u[i]] + ["%g" % sc.p[i], "%g" % sc.n[i]]
        lines.append(", ".join(vals))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("which", ["dambreak2d", "static", "dambreak3d"])
def test_xml_in_csv_out_round_trip(tmp_path, which):
    sc = {"dambreak2d": scenes.dambreak2d, "static": scenes.static_pressure, "dambreak3d": lambda: scenes.dambreak3d(2.4e-2)}[which]()
    rng = np.random.default_rng(7)
    sc.u[:] = rng.normal(size=sc.u.shape) * 1e-3          # exercise the number formats: small, negative, exponent forms
    sc.p[:] = rng.uniform(0, 3e3, size=sc.count)
    sc.n[:] = rng.uniform(0, 7, size=sc.count)
    sc.p[0] = 1e-7; sc.p[1] = 123456789.0; sc.n[2] = 0.0001; sc.n[3] = 100000.0; sc.n[4] = 1e6
    xml = scenes.write_xml(sc, str(tmp_path / "in.xml"), start_time=0.01, end_time=0.5, output_interval=5e-3)
    out = tmp_path / "result"
    out.mkdir()
    r = run(BIN3 if sc.env.dim == 3 else BIN2, "--check-io", xml, str(out))
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"Input XML file: {xml}" in r.stdout
    assert f"{sc.count} particles" in r.stdout                       # Main.cpp:180
    # outputIterationOffset = ceil(startTime / outputInterval) = 2 names the first file (Main.cpp:336)
    text = (out / "particles_00002.csv").read_text()
    assert text == expected_csv(sc)
    # the environment the driver derives (Main.cpp:234: maxDt = outputInterval / minStepCountPerOutput; Environment.hpp:129-161)
    env_line = next(l for l in r.stdout.splitlines() if l.startswith("environment:"))
    kv = dict(t.split("=") for t in env_line.split()[1:])
    assert int(kv["dim"]) == sc.env.dim
    assert float(kv["l_0"]) == sc.env.l0
    assert float(kv["MaxDx"]) == sc.env.courant * sc.env.l0
    assert float(kv["R_e"]) == sc.env.r_e_by_l0 * sc.env.l0
    # derived constants of the host-side Environment against the reference build / its CPU restatement, bit for bit
    from oracle import bind
    ref = (bind.RefComputer if bind.available(sc.env.dim, sc.env.central_gravity) else bind.PortComputer).from_scene(sc).env_values()
    for ours, theirs in (("n0", "n0"), ("MaxDx", "MaxDx"), ("R_e", "R_e"), ("NeighborLength", "NeighborLength")):
        assert float(kv[ours]) == ref[theirs], (ours, kv[ours], ref[theirs])
    # and what comes back is what went in, to the 6 significant digits of the format
    back = scenes.read_result_csv(str(out / "particles_00002.csv"))
    assert np.array_equal(back["type"], sc.type)
    assert np.allclose(back["x"], sc.x, rtol=1e-5, atol=1e-12)


CSV_OK = "Type, x, z, u, w, p, n\n0, 0.0, 0.0, 0, 0, 0, 0\n1, 0.008, 0.0, 0, 0, 0, 0\n"


def _xml(particles_text, ptype="csv", extra_env=""):
    return f"""<?xml version="1.0" encoding="utf-8"?>
<openmps>
  <condition>
    <startTime value="0" /> <!-- comment -->
    <endTime value="1.0" />
    <outputInterval value="0.005" />
    <eps value="1e-10" />
  </condition>
  <environment>
    <l_0 value="0.008" />
    <minStepCountPerOutput value="10" />
    <courant value="0.1" />
    <g value="9.8" /> <rho value="998.20" /> <nu value="1.004e-6" /> <r_eByl_0 value="2.4" />
    <surfaceRatio value="0.97" />
    <minX value="-0.032" /> <minY value="-0.000" /> <minZ value="-0.032" />
    <maxX value=" 0.584" /> <maxY value=" 0.000" /> <maxZ value=" 0.584" />
    <c value="1.5" /> --> <!-- stray text between elements, as in the reference's Sample.xml -->
    {extra_env}
  </environment>
  <particles type="{ptype}">
{particles_text}  </particles>
</openmps>
"""


def _check(tmp_path, text):
    f = tmp_path / "in.xml"
    f.write_text(text)
    (tmp_path / "result").mkdir(exist_ok=True)
    return run(BIN2, "--check-io", str(f), str(tmp_path / "result"))


def test_sample_like_layout_and_tolerances(tmp_path):
    # header-name driven columns in any order, blanks / tabs anywhere, empty lines skipped (Main.cpp:70-183)
    body = "\n\n  w ,u,\tn, p, z, x, Type\n\n 0, 0,0,0, 0.5, 0.25, 0 \n\n0,0,0,0,-0.008,\t0.016,2\n"
    r = _check(tmp_path, _xml(body))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "2 particles" in r.stdout
    assert (tmp_path / "result" / "particles_00000.csv").read_text() == "Type, x, z, u, w, p, n\n0, 0.25, 0.5, 0, 0, 0, 0\n2, 0.016, -0.008, 0, 0, 0, 0\n"
    # maxX value=" 0.584": leading blank accepted, as by the reference's stream extraction
    assert "MaxDt=0.0005" in r.stdout                      # 0.005 / 10 < sqrt(2 C l0 / g)


@pytest.mark.parametrize("body,message", [
    ("Type, x, z, u, w, p, n, q\n0, 0, 0, 0, 0, 0, 0, 0\n", "Illegal header item in input csv"),   # Main.cpp:128-131
    ("Type, x, z, u, w, p\n0, 0, 0, 0, 0, 0\n", "Some header item doesn't exist"),                 # Main.cpp:136-140
])
def test_reference_error_messages(tmp_path, body, message):
    r = _check(tmp_path, _xml(body))
    assert r.returncode != 0
    assert message in r.stderr


def test_particles_type_must_be_csv(tmp_path):
    r = _check(tmp_path, _xml(CSV_OK, ptype="binary"))
    assert r.returncode != 0 and "Not Implemented!" in r.stderr                                    # Main.cpp:196-198


def test_missing_value_and_bad_number(tmp_path):
    r = _check(tmp_path, _xml(CSV_OK).replace('<l_0 value="0.008" />', ""))
    assert r.returncode != 0 and "No such node" in r.stderr
    r = _check(tmp_path, _xml(CSV_OK).replace('<g value="9.8" />', '<g value="9.8 m/s2" />'))
    assert r.returncode != 0 and "conversion of data" in r.stderr


def test_progress_line_format():
    # "#%3$05d: t=%1$8.4lf (%2$05d), %10$12d particles, @ %4$02d/%5$02d %6$02d:%7$02d:%8$02d (%9$8.2lf)"  (Main.cpp:334)
    src = r'''
#include "DriverIo.hpp"
int main() { std::tm t{}; t.tm_mon = 9; t.tm_mday = 17; t.tm_hour = 8; t.tm_min = 5; t.tm_sec = 3;
  std::cout << OpenMps::DriverIo::ProgressLine(0.125, 257, 25, 1323, t, 12.3456) << std::endl; return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "t.cpp"), "w") as f:
            f.write(src)
        inc = os.path.join(ROOT, "include")
        subprocess.run(["g++", "-std=c++17", "-I" + os.path.join(inc, "openmps"), "-I" + inc, os.path.join(d, "t.cpp"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout
    assert out == "#00025: t=  0.1250 (00257),         1323 particles, @ 10/17 08:05:03 (   12.35)\n"


REF_SAMPLE = "/root/reference/Benchmark/Sample/Sample.xml"


@pytest.mark.skipif(not os.path.exists(REF_SAMPLE), reason="the reference tree is only mounted in the build container")
def test_reads_the_reference_sample_xml(tmp_path):
    """The shipped Benchmark/Sample/Sample.xml (BOM, comments in Japanese, a stray '-->' text node, unused elements) parses to
    the same 1 323 particles our generator makes, and the written CSV equals the %g text of those particles."""
    (tmp_path / "result").mkdir()
    r = run(BIN2, "--check-io", REF_SAMPLE, str(tmp_path / "result"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "1323 particles" in r.stdout
    sc = scenes.dambreak2d()
    assert (tmp_path / "result" / "particles_00000.csv").read_text() == expected_csv(sc)
    cond = next(l for l in r.stdout.splitlines() if l.startswith("condition:"))
    kv = dict(t.split("=") for t in cond.split()[1:])
    assert (float(kv["eps"]), float(kv["startTime"]), float(kv["endTime"]), float(kv["outputInterval"])) == (1e-10, 0.0, 1.0, 0.005)


def _model_input_from_csv(text):
    """Python restatement of InputFromCsv (Main.cpp:70-183): split at '\\n', leading zero-length lines skipped, blanks and tabs
    removed from every line, header names decide the columns, empty lines skipped."""
    lines = text.split("\n")
    k = 0
    while lines[k] == "":
        k += 1

    def items(line):
        s = line.replace(" ", "").replace("\t", "")
        return s.split(",") if s else []
    names = ["Type", "x", "z", "u", "w", "p", "n"]
    head = items(lines[k]); k += 1
    for h in head:
        if h not in names:
            raise ValueError("Illegal header item in input csv")
    if any(nm not in head for nm in names):
        raise ValueError("Some header item doesn't exist")
    col = {nm: head.index(nm) if head.count(nm) == 1 else max(i for i, h in enumerate(head) if h == nm) for nm in names}
    rows = []
    for line in lines[k:]:
        d = items(line)
        if d:
            rows.append((int(d[col["Type"]]), float(d[col["x"]]), float(d[col["z"]]), float(d[col["u"]]), float(d[col["w"]]), float(d[col["p"]]), float(d[col["n"]])))
    return rows


def test_csv_parser_against_a_model_of_the_reference_parser(tmp_path):
    rng = np.random.default_rng(20261017)
    names = ["Type", "x", "z", "u", "w", "p", "n"]

    def pad(tok):
        ws = ["", " ", "  ", "\t", " \t"]
        return ws[rng.integers(len(ws))] + tok + ws[rng.integers(len(ws))]

    def number():
        kind = rng.integers(5)
        v = [rng.normal(), rng.normal() * 1e-9, rng.normal() * 1e7, float(rng.integers(-5, 6)), 0.0][kind]
        return [repr(float(v)), "%g" % v, "%.3e" % v, "%d" % int(v) if float(v).is_integer() else "%r" % float(v)][rng.integers(4)]
    for case in range(12):
        order = list(rng.permutation(names))
        lines = [""] * int(rng.integers(0, 3))
        lines.append(",".join(pad(h) for h in order))
        for _ in range(int(rng.integers(1, 40))):
            if rng.random() < 0.15:
                lines.append(["", "   ", "\t"][rng.integers(3)])          # empty lines anywhere after the header
                continue
            vals = {nm: number() for nm in names}
            vals["Type"] = str(int(rng.integers(0, 4)))
            lines.append(",".join(pad(vals[h]) for h in order))
        body = "\n".join(lines) + "\n"
        want = _model_input_from_csv(body)
        d = tmp_path / f"c{case}"
        d.mkdir()
        r = _check(d, _xml(body))
        assert r.returncode == 0, r.stdout + r.stderr
        assert f"{len(want)} particles" in r.stdout
        got = (d / "result" / "particles_00000.csv").read_text().splitlines()[1:]
        exp = [", ".join([str(t)] + ["%g" % v for v in rest]) for (t, *rest) in want]
        assert got == exp, case
