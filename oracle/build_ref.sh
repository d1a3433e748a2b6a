#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference (Bigfoot71/PixelForge) from the sources
# where they lie under /root/reference into oracle/_ref/ (git-ignored, binaries only, no source copied
# into the repo).  Recipe mirrors the reference's CMakeLists.txt:5,23,45-48,65-72 in Release mode:
#   gcc -std=gnu99 -O3 -DNDEBUG -mavx2 -fopenmp on every .c found recursively under src/.
# NEVER add -march=native / -mfma / -ffast-math: GCC would contract mul+add pairs into FMAs and the
# pixels would change.
#
# Outputs:
#   oracle/_ref/libpf_ref.so        verbatim reference (all 128 PF_API symbols)
#   oracle/_ref/libpf_ref_bfix.so   reference with the one-token bilinear fix of SURVEY.md Q7
#                                   (src/internal/color.h:141  aV4 -> bV4), patched in a /tmp scratch
#                                   copy; only the binary lands here.  Verbatim bilinear is undefined
#                                   behaviour upstream (reads an uninitialised vector).
set -euo pipefail
REF=${PF_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF/src not present (GPU box?) - using prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
CFLAGS="-std=gnu99 -O3 -DNDEBUG -mavx2 -fopenmp -fPIC -DPF_BUILD_SHARED -w"

build_one() {  # $1 = source root, $2 = output .so
  local src="$1" out="$2" tmp
  tmp="$(mktemp -d /tmp/pfref_obj.XXXXXX)"
  local i=0 objs=()
  while IFS= read -r f; do
    i=$((i+1))
    gcc $CFLAGS -c "$f" -o "$tmp/o$i.o"
    objs+=("$tmp/o$i.o")
  done < <(find "$src/src" -name '*.c' | sort)
  gcc -shared -fopenmp -o "$out" "${objs[@]}" -lm
  rm -rf "$tmp"
}

build_one "$REF" "$OUT/libpf_ref.so"

SCR="$(mktemp -d /tmp/pfref_bfix.XXXXXX)"
cp -r "$REF/src" "$SCR/src"
# Q7: pfiColorLerpSmooth_simd unpacks `b` into aV4 and leaves bV4 uninitialised.
sed -i '141s/pfiColorSIMDToVecF_simd(aV4, b, 4)/pfiColorSIMDToVecF_simd(bV4, b, 4)/' "$SCR/src/internal/color.h"
grep -n 'pfiColorSIMDToVecF_simd(bV4, b, 4)' "$SCR/src/internal/color.h" >/dev/null
build_one "$SCR" "$OUT/libpf_ref_bfix.so"
rm -rf "$SCR"
ls -la "$OUT"
