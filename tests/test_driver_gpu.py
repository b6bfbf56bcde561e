"""The drop-in driver end to end on the GPU: XML in, result/particles_%05d.csv and progress lines out,
against the same run made through the C ABI from Python."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmps_b200 import capi, scenes  # noqa: E402

BIN2 = os.path.join(ROOT, "openmps_b200", "bin", "OpenMps")


def _csv_text(st, dim):
    head = "Type, x, z, u, w, p, n" if dim == 2 else "Type, x, y, z, u, v, w, p, n"
    lines = [head]
    for i in range(len(st["type"])):
        vals = [str(int(st["type"][i]))] + ["%g" % v for v in st["x"][i]] + ["%g" % v for v in st["u"][i]] + ["%g" % st["p"][i], "%g" % st["n"][i]]
        lines.append(", ".join(vals))
    return "\n".join(lines) + "\n"


@pytest.mark.gpu
def test_driver_outputs_match_the_c_abi_run(tmp_path, monkeypatch):
    if not os.path.exists(BIN2):
        pytest.fail("driver not built: openmps_b200/bin/OpenMps is missing")
    monkeypatch.setenv("MPS_CG_ADAPTIVE", "0")          # frozen CTA split: both runs are then bit-identical
    sc = scenes.dambreak2d()
    interval, n_out = 5e-3, 3
    xml = scenes.write_xml(sc, str(tmp_path / "in.xml"), start_time=0.0, end_time=interval * n_out, output_interval=interval,
                           min_step_count_per_output=10)
    r = subprocess.run([BIN2, xml], capture_output=True, text=True, cwd=str(tmp_path), timeout=600, env=dict(os.environ))
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.splitlines()
    assert out[0] == f"Input XML file: {xml}" and out[1] == f"{sc.count} particles" and out[-1] == "finished"
    prog = [l for l in out if l.startswith("#")]
    assert len(prog) == n_out + 1
    pat = re.compile(r"^#(\d{5}): t=\s*([0-9.]+) \((\d{5,})\),\s+(\d+) particles, @ \d\d/\d\d \d\d:\d\d:\d\d \(\s*[0-9.]+\)$")
    m = [pat.match(l) for l in prog]
    assert all(m), prog
    assert [int(x.group(1)) for x in m] == list(range(n_out + 1))
    assert all(int(x.group(4)) == sc.count for x in m)

    # the same run through the C ABI: while (T < next) ForwardTime(), Main.cpp:370-376
    g = capi.GpuComputer.from_scene(sc)
    assert (tmp_path / "result" / "particles_00000.csv").read_text() == _csv_text(g.state(), 2)
    t_next, steps = 0.0, 0
    for k in range(1, n_out + 1):
        t_next += interval
        steps += g.run_until(t_next)
        assert (tmp_path / "result" / f"particles_{k:05d}.csv").read_text() == _csv_text(g.state(), 2), f"output {k}"
        assert int(m[k].group(3)) == steps
        assert abs(float(m[k].group(2)) - g.time()[0]) < 1e-4
    back = scenes.read_result_csv(str(tmp_path / "result" / f"particles_{n_out:05d}.csv"))
    assert np.array_equal(back["type"], g.state()["type"])
