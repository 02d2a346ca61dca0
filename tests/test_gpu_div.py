"""The Kerr right-hand side forms its 14 IEEE quotients as div_by(x, d, div_rcp(d)) with the
reciprocal half shared between quotients over the same denominator (csrc/lp_internal.cuh).  The
claim is bit equality with `/`: tools/div_check.cu sweeps 1.5e10 operand pairs (random, all-ones
and power-of-two mantissas, exponents up to +-450, numerator 1) against __ddiv_rn on the GPU."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shared_reciprocal_division_is_ieee(native, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available on this box")
    exe = str(tmp_path / "div_check")
    build = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false", "-O3",
                            "-I" + os.path.join(ROOT, "include"),
                            "-I" + os.path.join(ROOT, "light_path_tracer_b200", "csrc"),
                            os.path.join(ROOT, "tools", "div_check.cu"), "-o", exe],
                           capture_output=True, text=True, timeout=300)
    assert build.returncode == 0, build.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(run.stdout)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = [l for l in run.stdout.splitlines() if l.startswith("exponents")]
    assert len(lines) == 3 and all(l.split("pattern:")[1].split() == ["0"] * 8 for l in lines)
