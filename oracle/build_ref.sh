#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (cvxopt) from the sources
# where they lie under /root/reference into oracle/_ref/ (git-ignored).  Only the
# dense base/blas/lapack/misc_solvers extensions are built, linked against the
# LP64 OpenBLAS that scipy bundles (no gfortran / system BLAS in this image);
# cholmod/umfpack/amd (SuiteSparse, absent) become stubs that raise.
# Nothing under cvxopt_b200/ may import oracle/_ref; tests, smoke() and the
# bench's CPU baseline arm may.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CVXOPT_REFERENCE:-/root/reference}"
SRC="$REF/src"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
if [ ! -d "$SRC/C" ]; then
  echo "build_ref: $SRC not present; keeping prebuilt oracle/_ref (if any)"; exit 0
fi
mkdir -p "$OUT/cvxopt" "$OUT/.build"
LIBDIR="$($PY -c "import scipy,os;print(os.path.realpath(os.path.join(os.path.dirname(scipy.__file__),'..','scipy.libs')))")"
LIB="$(ls "$LIBDIR"/libscipy_openblas-*.so | head -1)"
B="$OUT/.build"
# 1. map every Fortran extern the six C files use onto scipy_<name> (LP64 build)
grep -ohE "\b[a-z][a-z0-9]+_\(" "$SRC"/C/{base,dense,sparse,blas,lapack,misc_solvers}.c | tr -d '(' | sort -u > "$B/syms"
nm -D "$LIB" | awk '$2=="T"{print $3}' | sed -n 's/^scipy_//p' | sort -u | comm -12 "$B/syms" - \
  | awk '{print "#define "$1" scipy_"$1}' > "$B/scipy_redef.h"
INC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
EXT="$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")"
CF="-O2 -fPIC -shared -I$INC -I$SRC/C -include $B/scipy_redef.h -Wno-incompatible-pointer-types -Wno-implicit-function-declaration -w"
LD="-L$LIBDIR -l:$(basename "$LIB") -Wl,-rpath,$LIBDIR -lm"
gcc $CF "$SRC/C/base.c" "$SRC/C/dense.c" "$SRC/C/sparse.c" -o "$OUT/cvxopt/base$EXT" $LD
for m in blas lapack misc_solvers; do
  gcc $CF "$SRC/C/$m.c" -o "$OUT/cvxopt/$m$EXT" $LD
done
# 2. the reference's pure-python layer, installed as a package (what pip would do)
cp "$SRC"/python/*.py "$OUT/cvxopt/"
printf 'version="0+oracle"\nversion_tuple=(0,0,0)\n__version__=version\n' > "$OUT/cvxopt/_version.py"
for m in cholmod umfpack amd; do
cat > "$OUT/cvxopt/$m.py" <<PYEOF
# stub: SuiteSparse is not available in this image
options = {}
def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError("$m.%s: SuiteSparse not built in the oracle" % name)
    return _missing
PYEOF
done
echo "build_ref: OK -> $OUT (openblas: $LIB)"
