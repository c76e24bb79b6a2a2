"""Host-side numerics of device building blocks that are written __host__ __device__:
the fused fp64 softplus must stay within a few ulps of log(1+exp(x)) everywhere."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_log1pexp_accuracy(tmp_path):
    exe = str(tmp_path / "softplus_host")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "softplus_host.cu")])
    out = subprocess.check_output([exe]).decode().split()
    max_abs, max_rel, special_ok = float(out[0]), float(out[1]), int(out[3])
    assert max_abs < 4e-15      # half an ulp of the largest results (|x| ~ 40)
    assert max_rel < 1e-9       # relative accuracy where the value is > 3e-7 (x > -15)
    assert float(out[7]) < 4e-16  # absolute error for x <= 0 (values in (0, log 2])
    assert float(out[5]) < 1e-15 and int(out[6]) == 1   # fast_exp: relative error, special values
    assert float(out[4]) < 4e-15  # y*eta - log(1+e^eta) fused term, |eta| <= 20
    assert special_ok == 1      # NaN propagates, huge |x| and 0 are exact
