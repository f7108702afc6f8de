#!/bin/bash
# Compiles the reference's OWN simulation*.cpp, unmodified and where they lie under /root/reference, against the MKL-API shim
# (oracle/mkl_shim/mkl.h) into importable `simulation` extension modules under oracle/_ref/<task>/.
# TEST INFRASTRUCTURE ONLY.  The -D macros are the ones the reference's setupC.py passes (Q/setupC.py:55, H/setupC.py:49) with the
# default arguments of each task (arguments.py), formatted with repr() exactly like setupC.py does.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="/root/reference/implementation codes"
[ -d "$REF" ] || { echo "no /root/reference here: oracle/_ref not rebuilt"; exit 0; }
PY=${PYTHON:-python}
PYINC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
NPINC=$($PY -c "import numpy; print(numpy.get_include())")
CXX=${CXX:-g++}
FLAGS="-O2 -std=c++14 -fPIC -shared -w -DMKL_ILP64 -I$HERE/mkl_shim -I$PYINC -I$NPINC"
build() {  # name, source, macros...
  local name="$1" src="$2"; shift 2
  mkdir -p "$HERE/_ref/$name"
  $CXX $FLAGS "$@" -o "$HERE/_ref/$name/simulation.so" "$src"
  echo "built oracle/_ref/$name/simulation.so"
}
M() { $PY -c "from math import pi; print(repr($1))"; }
build quartic "$REF/quartic oscillator/simulation_quart.cpp" -DX_MAX=8.5 -DGRID_SIZE=0.1 -DMASS=$(M "1./pi") -DLAMBDA=$(M "0.04*pi") -DMOMENT=5 &
build inverted_quartic "$REF/inverted quartic oscillator/simulation_quart.cpp" -DX_MAX=13.0 -DGRID_SIZE=0.05 -DMASS=$(M "1./pi") -DLAMBDA=$(M "-0.01*pi") -DMOMENT=5 &
build harmonic "$REF/harmonic oscillator/simulation.cpp" -DN_MAX=70 -DOMEGA=$(M "1.0*pi") &
build inverted_harmonic "$REF/inverted harmonic oscillator/simulation_i.cpp" -DN_MAX=180 -DOMEGA=$(M "1.0*pi") &
wait
