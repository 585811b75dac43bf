"""The Fortran subroutine API shipped as source (roadsurf_b200/fortran/RoadSurf.f90, RoadSurfVariables.f90):
no Fortran compiler exists here, so these tests check what can be checked without one -- the 14 public
procedures of the reference's module RoadSurf (src/RoadSurf.f90:257-270) are all there with the same dummy
argument lists, the five interoperable types have the components of the C ABI in order, and every C symbol
the module binds is exported by libroadsurf_b200.so."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "roadsurf_b200", "fortran")

# procedure -> dummy arguments in order (src/RoadSurf.f90:9-252)
EXPECTED = {
    "connectfortran2carrays": ["inpointers", "modelinput", "outpointers", "modeloutput"],
    "initialization": ["modelinput", "insettings", "settings", "modeloutput", "atm", "surf", "inputparam",
                       "localparam", "coupling", "phy", "ground", "condparam"],
    "checkvalues": ["modelinput", "i", "settings", "surf", "localparam"],
    "couplingoperations1": ["i", "coupling", "surf", "settings", "ground", "modelinput", "cp", "localparam"],
    "relaxationoperations": ["i", "atm", "settings", "ground"],
    "setcurrentvalues": ["i", "modelinput", "atm", "settings", "surf", "coupling", "ground"],
    "balancemodelonestep": ["swi", "lwi", "phy", "ground", "surf", "atm", "settings", "coupling", "modelinput",
                            "inputidx", "condparam"],
    "saveoutput": ["modeloutput", "i", "surf"],
    "checkendcoupling": ["i", "settings", "coupling", "surf"],
    "precipitationtostorage": ["settings", "cp", "precphase", "atm", "surf"],
    "modradiationbysurroundings": ["modelinput", "inputparam", "localparam", "i"],
    "wearfactors": ["snow2icefac", "tph", "surf", "wearf"],
    "roadcond": ["maxpormms", "surf", "atm", "settings", "cp", "wearf"],
    "calcalbedo": ["albedo", "surf", "cp"],
}


def _subroutines(path):
    """name -> argument list of every (module) subroutine statement of a free-form Fortran file."""
    text = open(path, errors="replace").read()
    text = re.sub(r"!.*", "", text)                       # comments
    text = re.sub(r"&\s*\n\s*&?", " ", text)              # continuation lines
    out = {}
    for m in re.finditer(r"^\s*(?:module\s+)?subroutine\s+(\w+)\s*\(([^)]*)\)", text, re.I | re.M):
        out.setdefault(m.group(1).lower(), [a.strip().lower() for a in m.group(2).split(",")])
    return out


def test_the_14_public_procedures_keep_their_argument_lists():
    subs = _subroutines(os.path.join(SHIM, "RoadSurf.f90"))
    for name, args in EXPECTED.items():
        assert subs.get(name) == args, (name, subs.get(name))
    text = open(os.path.join(SHIM, "RoadSurf.f90")).read().lower()
    for name in EXPECTED:
        assert re.search(r"public\s*::[^\n]*\b%s\b" % name, text), f"{name} is not public"
    assert "lastvalues" in subs                               # external, called by Simulation.f90:105


def test_expected_lists_are_the_reference_interface():
    ref = os.environ.get("ROADSURF_REFERENCE_ROOT", "/root/reference")
    path = os.path.join(ref, "src", "RoadSurf.f90")
    if not os.path.exists(path):
        pytest.skip(f"reference tree not present at {ref}")
    subs = _subroutines(path)
    assert {k: subs[k] for k in EXPECTED} == EXPECTED
    public = re.findall(r"public\s*::\s*(\w+)", open(path).read(), re.I)
    assert sorted(p.lower() for p in public) == sorted(EXPECTED)


def test_interoperable_types_follow_the_c_abi():
    from roadsurf_b200 import abi
    text = re.sub(r"!.*", "", open(os.path.join(SHIM, "RoadSurfVariables.f90")).read())
    for cname, ctype in (("InputPointers", abi.InputPointers), ("OutputPointers", abi.OutputPointers),
                         ("InputSettings", abi.InputSettings), ("InputParameters", abi.InputParameters),
                         ("LocalParameters", abi.LocalParameters)):
        m = re.search(r"type\s*,\s*bind\(C\)\s*::\s*%s\b(.*?)end\s+type" % cname, text, re.I | re.S)
        assert m, cname
        comps = []
        for line in m.group(1).splitlines():
            if "::" in line:
                comps += [c.strip().split("=")[0].strip() for c in line.split("::", 1)[1].split(",")]
        want = [n for n, _ in ctype._fields_ if not n.startswith("_")]
        assert [c.lower() for c in comps] == [w.lower() for w in want], cname
    # the 16 type names of the reference's module (src/RoadSurfVariables.f90:13-28)
    for t in ("InputArrays", "OutputArrays", "PhysicalParameters", "GroundVariables", "SurfaceVariables", "AtmVariables",
              "CouplingVariables", "ModelSettings", "InputRadiationCoefficient", "RoadCondParameters", "WearingFactors"):
        assert re.search(r"type\s*::\s*%s\b" % t, text, re.I), t


def test_every_bound_c_symbol_is_exported():
    from roadsurf_b200 import build, lib
    build.build_library()
    handle = lib.load()
    names = set()
    for f in ("RoadSurf.f90", "roadsurf_b200.f90"):
        names |= set(re.findall(r'bind\(C,\s*name="(\w+)"\)', open(os.path.join(SHIM, f)).read()))
    assert {"roadsurf_session_open", "roadsurf_step", "roadsurf_session_fetch", "runsimulation"} <= names
    for n in names:
        assert hasattr(handle, n), n
