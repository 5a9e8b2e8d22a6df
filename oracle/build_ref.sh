#!/usr/bin/env bash
# Build the REFERENCE's own path functions into oracle/_ref/ (test infrastructure only).
#
# Inputs are read where they lie under /root/reference (never copied into the repo):
#   src/bin/evaluate_path_bidir_mala_<c>_<l>_static.c           forward value   (gcc -O3, src/chad.cpp:912)
#   src/bin/evaluate_path_bidir_mala_<c>_<l>_static_derv.ispc   reverse-mode gradient (bundled ispc 1.11,
#                                                               flags of src/chad.cpp:948)
#   src/bin/evaluate_path_bidir_<c>_<l>_static_derv.ispc        forward-mode gradient + Hessian (H2MC library)
# Outputs: oracle/_ref/libpathref_mala.so, oracle/_ref/libpathref_hess.so (git-ignored; they
# travel to the GPU box with the repo snapshot).  ISA pinned to AVX2 so the objects run on any
# x86-64 host, not just this container's CPU.
#   MAXLEN_MALA (default 8) / MAXLEN_HESS (default 5): largest c+l-1 to build.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
MAXLEN_MALA=${MAXLEN_MALA:-8}
MAXLEN_HESS=${MAXLEN_HESS:-5}
JOBS=${JOBS:-$(nproc)}
if [ ! -d "$REF/src/bin" ]; then echo "reference not present at $REF: keeping prebuilt oracle/_ref" >&2; exit 0; fi
mkdir -p "$OBJ"
ISPC="$REF/ispc/bin/ispc"
CC=/usr/bin/gcc
jobs_file="$OBJ/jobs.txt"; : > "$jobs_file"
mala_objs=(); hess_objs=()
for c in 1 2 3 4 5 6 7 8 9; do for l in 0 1 2 3 4 5 6 7 8; do
  len=$((c + l - 1))
  if [ $((c + l)) -le 2 ]; then continue; fi
  if [ $len -le $MAXLEN_MALA ]; then
    n="evaluate_path_bidir_mala_${c}_${l}_static"
    [ -f "$OBJ/$n.o" ] || echo "$CC -O3 -c -fPIC -o $OBJ/$n.o $REF/src/bin/$n.c" >> "$jobs_file"
    [ -f "$OBJ/${n}_derv.o" ] || echo "$ISPC -O3 --math-lib=default --opt=fast-math --woff --pic --target=avx2-i32x8 -o $OBJ/${n}_derv.o $REF/src/bin/${n}_derv.ispc" >> "$jobs_file"
    mala_objs+=("$OBJ/$n.o" "$OBJ/${n}_derv.o")
  fi
  if [ $len -le $MAXLEN_HESS ]; then
    n="evaluate_path_bidir_${c}_${l}_static"
    [ -f "$OBJ/${n}_derv.o" ] || echo "$ISPC -O3 --math-lib=default --opt=fast-math --woff --pic --target=avx2-i32x8 -o $OBJ/${n}_derv.o $REF/src/bin/${n}_derv.ispc" >> "$jobs_file"
    hess_objs+=("$OBJ/${n}_derv.o")
  fi
done; done
echo "compiling $(wc -l < "$jobs_file") objects with $JOBS jobs"
xargs -P "$JOBS" -I{} sh -c '{}' < "$jobs_file"
$CC -shared -fPIC -o "$OUT/libpathref_mala.so" "${mala_objs[@]}" -lm
$CC -shared -fPIC -o "$OUT/libpathref_hess.so" "${hess_objs[@]}" -lm
ls -la "$OUT"/*.so
