"""The C ABI consumed from C (`-m gpu`): tests/cabi/q2_cabi.c includes include/flockgpu.h as C11, links
flock_b200/libflockgpu.so and runs NEXMark q2 through flock_context_* on hand-built ArrowArrays -- no Python, no ctypes
prototypes in between.  The CPU half (the header is valid C, the program links against every symbol it uses) runs in
tests/test_host.py."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
CABI = ROOT / "tests" / "cabi"


def build_cabi_program(out: Path) -> None:
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(ROOT / "include"), str(CABI / "q2_cabi.c"), "-L", str(ROOT / "flock_b200"),
           "-lflockgpu", f"-Wl,-rpath,{ROOT / 'flock_b200'}", "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


@pytest.mark.gpu
def test_q2_through_the_c_abi_from_c(tmp_path):
    exe = tmp_path / "q2_cabi"
    build_cabi_program(exe)
    r = subprocess.run([str(exe), str(CABI / "q2_plan.json")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "bit-exact" in r.stdout
