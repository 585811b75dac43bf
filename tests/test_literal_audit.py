"""scripts/audit_literals.py as a test: every REAL literal of the reference's hot-path Fortran occurs,
with the value a Fortran compiler gives it (REAL(4) unless D-exponent / kind suffix), in the oracle
function that cites the file, and the kernel uses no float32 literal the reference does not have.
Runs against the committed manifest (tests/golden/fortran_literals.json: numbers only); where the
reference tree exists the manifest itself is checked against it."""
import json
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import audit_literals as al  # noqa: E402


@pytest.fixture(scope="module")
def manifest():
    return json.load(open(al.MANIFEST))


def test_audit_is_clean(manifest):
    findings, _, _ = al.audit(manifest)
    assert not findings, "\n".join(findings)


def test_manifest_matches_reference_tree(manifest):
    ref = os.environ.get("ROADSURF_REFERENCE_ROOT", "/root/reference")
    if not os.path.exists(os.path.join(ref, "src", "RoadSurf.f90")):
        pytest.skip(f"reference tree not present at {ref}: auditing against the committed manifest only")
    fresh = json.loads(json.dumps(al.build_manifest(ref)))
    assert fresh == manifest, "tests/golden/fortran_literals.json is stale: scripts/audit_literals.py --write-manifest"


@pytest.mark.parametrize("old,new,what", [
    ("F4(0.61078)", "F4(0.61087)", "mistyped digit"),
    ("F4(273.15)", "R(273.15)", "double instead of REAL(4)"),
    ("F4(17.269)", "F4(17.296)", "mistyped digit"),
    ("F4(360.98564736629)", "R(360.98564736629)", "double instead of REAL(4)"),
])
def test_audit_detects_a_transcription_error_in_the_oracle(manifest, tmp_path, old, new, what):
    src = open(os.path.join(ROOT, "oracle", "roadsurf_oracle.hpp")).read()
    assert old in src
    mutated = tmp_path / "roadsurf_oracle.hpp"
    mutated.write_text(src.replace(old, new))
    findings, _, _ = al.audit(manifest, oracle_path=str(mutated))
    assert findings, f"audit missed: {what} ({old} -> {new})"


def test_audit_detects_a_wrong_literal_in_the_kernel(manifest, tmp_path):
    src = open(os.path.join(ROOT, "roadsurf_b200", "csrc", "rs_kernel.cu")).read()
    assert "F4(0.47496)" in src
    mutated = tmp_path / "rs_kernel.cu"
    mutated.write_text(src.replace("F4(0.47496)", "F4(0.47469)"))
    findings, _, _ = al.audit(manifest, kernel_paths=(str(mutated),))
    assert any("0.47469" in f for f in findings)


def test_real4_rounding_of_the_known_hard_cases(manifest):
    """The literals SURVEY.md section 7.1 singles out."""
    vals = {e["value"] for entries in manifest.values() for e in entries}
    assert 360.98565673828125 in vals          # 360.98564736629 as REAL(4)
    assert 273.149993896484375 in vals         # 273.15
    assert 1524.5 in vals                      # 1.5245D3 stays exact
    sun = {e["value"]: e["real4"] for e in manifest["src/SunPosition.f90"]}
    assert sun[1524.5] is False
