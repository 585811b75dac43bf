"""The kernel's exp / log are the host libm's algorithm (roadsurf_b200/csrc/rs_libm.h: glibc's table-driven
exp and log, operation order of its FMA code path).  Compiled for the host, that code must reproduce
the libm of this process bit for bit; the CUDA build executes the same IEEE operations."""
import numpy as np

from roadsurf_b200 import lib


def test_host_build_of_the_kernels_exp_and_log_equals_libm_bit_for_bit():
    bad = lib.selftest_libm(6_000_000, seed=11)
    assert bad == [0, 0], bad


def test_tables_come_from_this_systems_libm():
    """The committed table header equals what scripts/gen_libm_tables.py extracts here."""
    import os, subprocess, sys, tempfile, shutil
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "roadsurf_b200", "csrc", "rs_libm_tables.h")
    committed = open(hdr).read()
    with tempfile.TemporaryDirectory() as tmp:
        shutil.copytree(os.path.join(root, "scripts"), os.path.join(tmp, "scripts"))
        os.makedirs(os.path.join(tmp, "roadsurf_b200", "csrc"))
        subprocess.run([sys.executable, os.path.join(tmp, "scripts", "gen_libm_tables.py")], check=True,
                       capture_output=True)
        fresh = open(os.path.join(tmp, "roadsurf_b200", "csrc", "rs_libm_tables.h")).read()
    assert fresh == committed
