#!/usr/bin/env python
"""Mechanical audit of numeric literals: reference Fortran -> oracle restatement -> CUDA kernel.

The commonest transcription error when restating the reference without a Fortran compiler is a real
literal with the wrong precision or a mistyped digit: an un-suffixed Fortran literal is REAL(4)
(`273.15` is 273.149993896484375), only `1.5245D3` / `1.0_8` forms are REAL(8).  This script

  1. extracts every REAL literal of the 13 hot-path Fortran files (comments, strings and I/O
     statements stripped), with the value the Fortran compiler gives it (float32-rounded unless it
     has a D exponent or a kind suffix), per file;
  2. extracts every literal of oracle/roadsurf_oracle.hpp (F4(x) and x##f forms are float32-rounded,
     everything else is a double), attributed to the reference file its nearest preceding
     `src/X.f90` / `Simulation.f90` citation names;
  3. requires every Fortran literal VALUE of a file to occur among the oracle literals attributed
     to that file (allow-listed exceptions: dead code the oracle does not restate, each justified
     below), and every float32-form literal of the oracle to be a REAL(4) literal of that file;
  4. does the same for the CUDA kernel (roadsurf_b200/csrc/rs_kernel.cu, rs_model.cpp) against the
     union of the Fortran files.

    python scripts/audit_literals.py [--reference /root/reference] [--write-manifest]

With the reference tree present the Fortran side is read from it (and `--write-manifest` refreshes
tests/golden/fortran_literals.json: numbers only, no reference source); without it the committed
manifest is used, so the audit also runs where /root/reference does not exist.
Exit code 0 = clean, 1 = findings (printed).
"""
import argparse
import json
import os
import re
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MANIFEST = os.path.join(ROOT, "tests", "golden", "fortran_literals.json")
FORTRAN_FILES = ["src/BalanceModel.f90", "src/BoundaryLayer.f90", "src/Cond.f90", "src/ConnectFortran2Carrays.f90",
                 "src/Coupling.f90", "src/Initialization.f90", "src/InputOutput.f90", "src/ModRadiation.f90",
                 "src/Relaxation.f90", "src/Storage.f90", "src/SunPosition.f90", "src/RoadSurf.f90",
                 "examples/example1/src/Simulation.f90"]


def f32(x):
    return struct.unpack("f", struct.pack("f", float(x)))[0]


NUM = re.compile(r"(?<![\w.])(\d+\.\d*|\.\d+|\d+(?=[dDeE][+-]?\d))([dDeE][+-]?\d+)?(_\w+)?")


def strip_fortran(line):
    """Drop string literals and the trailing comment of one Fortran source line."""
    out, q = [], None
    for ch in line:
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
            continue
        if ch == "!":
            break
        out.append(ch)
    return "".join(out)


def fortran_literals(path):
    """[(line number, text, value, is_real4)] of the REAL literals of a Fortran file."""
    res = []
    for ln, raw in enumerate(open(path, errors="replace"), 1):
        line = strip_fortran(raw)
        low = line.strip().lower()
        if low.startswith(("write", "print", "format", "#")):
            continue
        for m in NUM.finditer(line):
            mant, expo, kind = m.group(1), m.group(2) or "", m.group(3) or ""
            # `1.and.` style operators never occur in the reference; guard against `1.e` of `.eq.`
            tail = line[m.end():m.end() + 3].lower()
            if mant.endswith(".") and not expo and re.match(r"(and|or\.|eq\.|ne\.|lt\.|gt\.|le\.|ge\.|not)", tail):
                continue
            dbl = bool(kind) or (expo[:1] in "dD" and expo != "")
            txt = mant + expo.replace("d", "e").replace("D", "e")
            val = float(txt)
            res.append((ln, m.group(0), val if dbl else f32(val), not dbl))
    return res


CNUM = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(f)?")
CITE = re.compile(r"(src/\w+\.f90|examples/example\d/src/Simulation\.f90|Simulation\.f90)")


def c_literals(path, attribute=True):
    """[(line, text, value, is_float32_form, attributed Fortran file or None)] of a C++/CUDA source."""
    res, cur, section = [], None, None
    in_block = False
    for ln, raw in enumerate(open(path), 1):
        line = raw
        cites = CITE.findall(line) if attribute else []
        if cites and "//" in line:
            cur = tuple("examples/example1/src/Simulation.f90" if c.endswith("Simulation.f90") else c for c in cites)
            if "====" in line:
                section = cur          # a section banner: `// ==== src/X.f90 ====`
        elif attribute and re.match(r"\s*// :\d+", line):
            cur = section              # a relative citation `// :a-b` refers to the section's file
        # strip comments and strings
        if in_block:
            if "*/" in line:
                line = line.split("*/", 1)[1]
                in_block = False
            else:
                continue
        line = re.sub(r'"(\\.|[^"\\])*"', '""', line)
        line = re.sub(r"/\*.*?\*/", "", line)
        i_line, i_block = line.find("//"), line.find("/*")
        if i_block >= 0 and (i_line < 0 or i_block < i_line):
            line = line[:i_block]
            in_block = True
        elif i_line >= 0:
            line = line[:i_line]
        if line.lstrip().startswith("#"):
            continue
        # F4(x): float32-rounded
        spans = []
        for m in re.finditer(r"F4\(\s*(-?)\s*([0-9.eE+-]+?)\s*\)", line):
            res.append((ln, m.group(0), f32(float(m.group(2))), True, cur))
            spans.append(m.span())
        for m in CNUM.finditer(line):
            if any(a <= m.start() < b for a, b in spans):
                continue
            val = float(m.group(1))
            is32 = m.group(2) == "f"
            res.append((ln, m.group(0), f32(val) if is32 else val, is32, cur))
    return res


# Fortran literals that the restatement legitimately does not carry, per file: (value, why)
ALLOW_MISSING = {
    # the sky-view test `sky_view < 1.0 .and. sky_view > -0.01` (Simulation.f90:154) is restated once as
    # Model::sky_view_active next to its other user (CouplingOperations1), under the InputOutput banner
    "examples/example1/src/Simulation.f90": [(f32(0.01), "sky_view_active"), (1.0, "sky_view_active")],
    # MissValR = -9999.9 of CalcRh (:220): CalcRh is never called by the library;
    # the oracle restates it as a free function outside the banner sections
    "src/InputOutput.f90": [(f32(9999.9), "CalcRh, unused")],
}
# literal values that are only loop/format artefacts on the C side (not physics): ignored in the
# "oracle literal must exist in Fortran" direction
C_ONLY_OK = {0.0, 1.0, 2.0, 0.5, 1e-9}


def build_manifest(ref_root):
    man = {}
    for rel in FORTRAN_FILES:
        lits = fortran_literals(os.path.join(ref_root, rel))
        vals = {}
        for ln, txt, val, r4 in lits:
            key = repr(val)
            e = vals.setdefault(key, {"value": val, "real4": r4, "count": 0, "lines": []})
            e["count"] += 1
            e["real4"] = e["real4"] and r4
            if len(e["lines"]) < 12:
                e["lines"].append(ln)
        man[rel] = sorted(vals.values(), key=lambda e: e["value"])
    return man


def audit(man, verbose=False, oracle_path=None, kernel_paths=None):
    """Findings (strings) of the audit; oracle_path / kernel_paths let the tests audit mutated copies."""
    findings = []
    oracle = c_literals(oracle_path or os.path.join(ROOT, "oracle", "roadsurf_oracle.hpp"))
    by_file = {}
    for ln, txt, val, is32, cur in oracle:
        for c in (cur or (None,)):   # a comment citing two files attributes what follows to both
            by_file.setdefault(c, []).append((ln, txt, val, is32, cur))
    all_oracle_vals = {abs(v) for _, _, v, _, _ in oracle}
    union_fortran = {}
    for rel, entries in man.items():
        have = {abs(v) for _, _, v, _, _ in by_file.get(rel, [])}
        allow = {abs(v) for v, _ in ALLOW_MISSING.get(rel, [])}
        for e in entries:
            v = abs(e["value"])
            union_fortran.setdefault(v, e["real4"])
            if v in have or v in allow:
                continue
            where = "elsewhere in the oracle" if v in all_oracle_vals else "NOWHERE in the oracle"
            findings.append(f"{rel}: literal {e['value']!r} (lines {e['lines']}, real4={e['real4']}) not among the oracle "
                            f"literals attributed to this file; present {where}")
        # reverse direction: float32-form oracle literals must be REAL(4) literals of the file
        fvals = {abs(e["value"]): e["real4"] for e in entries}
        for ln, txt, val, is32, cur in by_file.get(rel, []):
            v = abs(val)
            if len(cur or ()) > 1 and any(v in {abs(e["value"]) for e in man.get(c, [])} for c in cur):
                continue             # belongs to the other file the comment cites
            if is32 and v not in fvals and v not in C_ONLY_OK:
                findings.append(f"oracle/roadsurf_oracle.hpp:{ln}: float32 literal {txt} ({val!r}) has no REAL(4) "
                                f"counterpart in {rel}")
            if not is32 and v in fvals and fvals[v] and f32(v) != v:
                findings.append(f"oracle/roadsurf_oracle.hpp:{ln}: literal {txt} is a double but {rel} has it as REAL(4)")
    # ---- the CUDA side: every literal of the kernel must be a literal value of the reference
    for rel in (kernel_paths or ("roadsurf_b200/csrc/rs_kernel.cu", "roadsurf_b200/csrc/rs_model.cpp")):
        for ln, txt, val, is32, _ in c_literals(os.path.join(ROOT, rel), attribute=False):
            v = abs(val)
            if is32 and v not in union_fortran and v not in C_ONLY_OK:
                findings.append(f"{rel}:{ln}: float32 literal {txt} ({val!r}) is not a REAL(4) literal of the reference")
            if not is32 and v in union_fortran and union_fortran[v] and f32(v) != v:
                findings.append(f"{rel}:{ln}: literal {txt} is a double but the reference has it as REAL(4)")
    # every inexact REAL(4) literal of the hot path must appear float32-rounded in the kernel sources
    kvals = set()
    for rel in ("roadsurf_b200/csrc/rs_kernel.cu", "roadsurf_b200/csrc/rs_model.cpp", "roadsurf_b200/csrc/rs_libm.h"):
        kvals |= {abs(v) for _, _, v, _, _ in c_literals(os.path.join(ROOT, rel), attribute=False)}
    return findings, kvals, union_fortran


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("ROADSURF_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--write-manifest", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    args = ap.parse_args()
    have_ref = os.path.exists(os.path.join(args.reference, "src", "RoadSurf.f90"))
    if have_ref:
        man = build_manifest(args.reference)
        if args.write_manifest:
            with open(MANIFEST, "w") as f:
                json.dump(man, f, indent=0, sort_keys=True)
            print("wrote", MANIFEST)
    else:
        man = json.load(open(MANIFEST))
    findings, kvals, union = audit(man, args.verbose)
    n = sum(len(v) for v in man.values())
    print(f"audit_literals: {n} distinct (file, value) REAL literals in {len(man)} Fortran files "
          f"({'reference tree' if have_ref else 'committed manifest'}); {len(findings)} finding(s)")
    for f in findings:
        print("  " + f)
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
