"""SURVEY §8(b) — the boundary is a DROP-IN: the reference's own, unmodified Python model code
(pytorch/model/*.py) runs on top of this repo's operators and reproduces the golden vectors made from the same
code on the CPU oracle.  Three ways in (tests/dropin_runner.py): through the Python operator API, through the
reference's native ABI (its own pointops.py + its own pybind glue linked against libcbops.so's *_cuda_launcher
symbols), and with this repo's Loss / ContrastHead taking the reference's constructor and forward arguments.
Skipped when baseline/_ref (a build-time copy of the reference's .py files, never committed) is absent."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFPY = os.path.join(ROOT, "baseline", "_ref", "pytorch", "model", "blocks.py")
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin", "pointops_cuda.so")


def _run(mode):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_runner.py"), mode], capture_output=True, text=True,
                       timeout=900)
    print(r.stdout[-3000:])
    lines = [l for l in r.stdout.splitlines() if l.startswith("DROPIN ")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-4000:])
    return json.loads(lines[-1][7:])


def _check(res):
    assert res["logits_err"] < 1e-4, res
    for a, b in zip(res["loss"], res["loss_ref"]):
        assert abs(a - b) <= 1e-4 * abs(b) + 1e-7, res
    assert max(res["latent_err"]) < 5e-4, res
    assert not res["grad_failures"], res
    assert any("libcbops.so" in l for l in res["loaded"]), res


@pytest.mark.skipif(not os.path.exists(REFPY), reason="baseline/_ref (copy of the reference's model code) not built")
@pytest.mark.parametrize("mode", ["python_api", "loss_adapter"])
def test_reference_model_code_on_cbops_python_api(mode):
    _check(_run(mode))


@pytest.mark.skipif(not (os.path.exists(REFPY) and os.path.exists(DROPIN)), reason="drop-in build of the reference glue absent")
def test_reference_model_code_on_cbops_native_abi():
    res = _run("native_abi")
    _check(res)
    assert any("dropin" in l for l in res["loaded"]) and not any(l.endswith("_ref/pointops_cuda.so") for l in res["loaded"]), res
