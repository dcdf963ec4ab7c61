"""GPU test: the C++ host mirror of the reference interface (quisquis-rust_b200/host/quisquis.hpp) compiled against
libqq_b200.so and run on fixtures whose expected bytes come from the oracle."""
import os
import subprocess

import pytest

import ristretto_ref as R
from qq_testlib import Stream, make_account, sb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror(tmp_path, pkg):
    st = Stream(b"cpp-mirror")
    rec = b""
    for i in range(9):
        v = [0, 5, 0, 3, 7, 0, 0, 1, 2][i]
        acc, sk, _ = make_account(st, v)
        bl, u, c = sb(st.scalar() % 2**40), st.scalar_bytes(), st.scalar_bytes()
        exp, es = R.update_account(acc, bl, u, c)
        assert es == 0
        rec += acc + sb(sk) + sb(v) + bl + u + c + exp
    fx = tmp_path / "fixture.bin"
    fx.write_bytes(rec)
    # second fixture: a destroy-account proof and a dark-tx proof (reference scenarios verifier.rs:1455-1479, :1075-1111)
    from qq_testlib import scenario_dark_tx, scenario_destroy
    accs, z, x = scenario_destroy(st, 4)
    sg = b"".join(accs) + b"".join(sb(v) for v in z) + sb(x)
    d, o, z, x = scenario_dark_tx(st, 4)
    sg += b"".join(d) + b"".join(o) + sb(z[0]) + sb(z[1]) + sb(x)
    fs = tmp_path / "sigma.bin"
    fs.write_bytes(sg)
    # third fixture: sender-account proof + aggregated range proof on one transcript (verifier.rs:1525-1628)
    from qq_testlib import scenario_range_batch
    k = scenario_range_batch(st)
    rg = b"".join(k[0]) + b"".join(k[1]) + k[2] + b"".join(sb(v) for v in k[3] + k[4] + k[5]) + sb(k[6]) + b"".join(k[7]) + k[8]
    fr = tmp_path / "range.bin"
    fr.write_bytes(rg)
    exe = tmp_path / "host_api_test"
    libdir = os.path.join(ROOT, "quisquis-rust_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "host_api_test.cpp"),
                           "-L" + libdir, "-lqq_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe), str(fx), str(fs), str(fr)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "HOST_API_TEST OK" in out.stdout, out.stdout + out.stderr


def test_cpp_multi_device_handle(tmp_path, pkg):
    """qq_init_multi / qq_multi_* through the C ABI from plain C++ (tests/csrc/multi_api_test.cpp): account batches, commitments
    and ONE MSM split over every visible GPU equal the single-context results byte for byte."""
    exe = tmp_path / "multi_api_test"
    libdir = os.path.join(ROOT, "quisquis-rust_b200")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "csrc", "multi_api_test.cpp"),
                           "-L" + libdir, "-lqq_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_API_TEST OK" in out.stdout, out.stdout + out.stderr
