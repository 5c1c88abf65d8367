"""The README quick start must keep running as written (examples/quickstart.py is generated from it)."""
import os
import runpy

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_readme_snippet_and_example_are_the_same_text():
    readme = open(os.path.join(ROOT, "README.md")).read()
    i = readme.index("```python\nimport torch, rocnrdma_b200 as rn")
    code = readme[i + len("```python\n"):readme.index("```", i + 10)]
    assert code in open(os.path.join(ROOT, "examples", "quickstart.py")).read()


@pytest.mark.gpu
def test_quickstart_runs(capsys):
    runpy.run_path(os.path.join(ROOT, "examples", "quickstart.py"), run_name="__main__")
    out = capsys.readouterr().out
    assert "GB/s device-timed" in out and "us per message" in out and "TFLOP/s including delivery" in out


@pytest.mark.gpu
@pytest.mark.parametrize("name,needle", [("verbs_wire.py", "GPU-posted RDMA write"), ("receive_and_multiply.py", "TFLOP/s in the kernel")])
def test_other_examples_run(name, needle):
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", name)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert needle in r.stdout
