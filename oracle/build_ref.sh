#!/bin/sh
# Builds the UNMODIFIED reference (fmidev/RoadSurf) into oracle/_ref/libroadsurf_ref.so:
# the library sources src/*.f90 plus the example's time loop examples/example1/src/Simulation.f90
# (the `runsimulation` BIND(C) entry), compiled from where they lie under the reference tree with
# the reference's own flag set (reference Makefile:28-37: -O2 -Ofast + MATHFLAGS).  Outputs go to
# oracle/_ref/ only (git-ignored; it travels to the GPU box with the snapshot).  No reference source
# is copied into this repository.
#
#   oracle/build_ref.sh [REFERENCE_ROOT]      default REFERENCE_ROOT=/root/reference
#   FC=gfortran (default), or any Fortran 2008 compiler with submodule support on PATH
#   REF_FLAVOUR=fast (default: the reference's flags) | strict (-O2, no fast-math: the bit-level pin)
#
# Exit codes: 0 built, 3 no Fortran compiler (the normal case in this image: gfortran, flang,
# nvfortran, ifx, lfortran and f2c are all absent, and gcc has no f951), 4 no reference tree,
# 1 compile error.  pyoracle.load_ref() and tests/test_oracle_ref.py call this script and report a
# skip WITH the reason on 3 / 4.
set -u
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${1:-${ROADSURF_REFERENCE_ROOT:-/root/reference}}
FC=${FC:-gfortran}
FLAVOUR=${REF_FLAVOUR:-fast}
OUT="$HERE/_ref"
if ! command -v "$FC" >/dev/null 2>&1; then
  echo "build_ref: no Fortran compiler ('$FC' not on PATH)" >&2
  exit 3
fi
if [ ! -f "$REF/src/RoadSurf.f90" ] || [ ! -f "$REF/examples/example1/src/Simulation.f90" ]; then
  echo "build_ref: reference tree not found under '$REF'" >&2
  exit 4
fi
MATHFLAGS="-funsafe-math-optimizations -fno-math-errno -fno-rounding-math -fno-trapping-math -fno-signed-zeros -freciprocal-math"
if [ "$FLAVOUR" = strict ]; then
  OPT="-O2 -ffp-contract=off"
  LIB="$OUT/libroadsurf_ref_strict.so"
else
  OPT="-O2 -Ofast $MATHFLAGS"
  LIB="$OUT/libroadsurf_ref.so"
fi
OBJ="$OUT/obj_$FLAVOUR"
mkdir -p "$OBJ" || exit 1
FLAGS="-fno-omit-frame-pointer -J$OBJ -I$REF/src -cpp -fPIC $OPT"
# module order of the reference Makefile:79-90: the types, the interface module, then its submodules
set -e
"$FC" $FLAGS -c "$REF/src/RoadSurfVariables.f90" -o "$OBJ/RoadSurfVariables.o"
"$FC" $FLAGS -c "$REF/src/RoadSurf.f90" -o "$OBJ/RoadSurf.o"
OBJS="$OBJ/RoadSurfVariables.o $OBJ/RoadSurf.o"
for f in BalanceModel BoundaryLayer Cond ConnectFortran2Carrays Coupling Initialization InputOutput \
         ModRadiation Relaxation Storage SunPosition; do
  "$FC" $FLAGS -c "$REF/src/$f.f90" -o "$OBJ/$f.o"
  OBJS="$OBJS $OBJ/$f.o"
done
"$FC" $FLAGS -c "$REF/examples/example1/src/Simulation.f90" -o "$OBJ/Simulation.o"
"$FC" -shared -rdynamic $OBJS "$OBJ/Simulation.o" -o "$LIB"
echo "build_ref: built $LIB ($FC $OPT)"
