#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the *unmodified reference* (G0T-cha/PysicalBasedRaytracer,
# read from $REF, default /root/reference) into oracle/_ref/libpbr_ref.so together with our own
# harness (oracle/ref_harness.cpp).  Reference sources are never copied into the repo: they are
# staged in a temp dir, compiled, and the temp dir is removed.  Outputs go only to oracle/_ref/.
#
# Portability edits applied to the staged copy (none changes arithmetic):
#   * GBK -> UTF-8 transcoding of comments
#   * '\' -> '/' and case fixes inside #include lines
#   * force-included shim: std headers, __declspec(x) -> nothing, _BitScanReverse via clz,
#     forward declaration of PBR::Clamp (MSVC-style late lookup in Core/PBR.h:180)
#   * Main/main.cpp and Shape/ModelLoad.* (assimp) are not compiled
#   * Core/Interaction.cpp:34-37: the explicit bsdf->~BSDF() (double destruction -> crash under
#     libstdc++) is removed.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
JOBS="${JOBS:-$(nproc)}"
if [ ! -d "$REF/Integrator" ]; then
  echo "build_ref: reference not present at $REF (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
STAGE="$(mktemp -d /tmp/pbr_ref_stage.XXXXXX)"
trap 'rm -rf "$STAGE"' EXIT
for d in Accelerator Camera Core Integrator Light Material Media Sampler Shape Texture include; do
  cp -r "$REF/$d" "$STAGE/$d"
done
chmod -R u+w "$STAGE"
rm -f "$STAGE/Shape/ModelLoad.cpp" "$STAGE/Shape/ModelLoad.h"
# transcode + include fixes
find "$STAGE" -type f \( -name '*.h' -o -name '*.cpp' \) ! -path "$STAGE/include/*" | while read -r f; do
  if ! iconv -f UTF-8 -t UTF-8 "$f" >/dev/null 2>&1; then
    iconv -f GBK -t UTF-8 -c "$f" > "$f.tmp" && mv "$f.tmp" "$f"
  fi
  sed -i -E '/^[[:space:]]*#[[:space:]]*include/ { s#\\#/#g;
      s#Core/interaction\.h#Core/Interaction.h#; s#Core/primitive\.h#Core/Primitive.h#;
      s#Core/frameBuffer\.h#Core/FrameBuffer.h#; s#Camera/orthographic\.h#Camera/Orthographic.h#;
      s#Sampler/clockRand\.h#Sampler/ClockRand.h#; s#"core/Spectrum\.h"#"Core/Spectrum.h"#;
      s#"include/stb_image\.h"#"stb_image.h"#; }' "$f"
done
# the one behavioural patch (crash fix, no arithmetic)
sed -i -E 's#^([[:space:]]*)bsdf->~BSDF\(\);#\1/* oracle: explicit dtor removed (double free under libstdc++) */;#' "$STAGE/Core/Interaction.cpp"
grep -q 'explicit dtor removed' "$STAGE/Core/Interaction.cpp" || { echo "patch failed" >&2; exit 1; }
cat > "$STAGE/shim.h" <<'SHIM'
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <memory>
#include <string>
#include <algorithm>
#include <vector>
#include <limits>
#include <iostream>
#include <unistd.h>
#define __declspec(x)
// MSVC resolves unqualified dependent names at instantiation; GCC needs the declaration first (PBR.h:180 uses Clamp before :184)
namespace PBR { template <typename T, typename U, typename V> inline T Clamp(T val, U low, V high); }
static inline unsigned char _BitScanReverse(unsigned long* idx, unsigned long v) {
  if (!v) return 0; *idx = 63 - __builtin_clzl(v); return 1; }
static inline unsigned char _BitScanReverse64(unsigned long* idx, unsigned long long v) {
  if (!v) return 0; *idx = 63 - __builtin_clzll(v); return 1; }
SHIM
CXXFLAGS="-std=c++14 -O2 -fopenmp -ffp-contract=off -fPIC -fpermissive -w -I$STAGE -I$STAGE/include -include $STAGE/shim.h"
cd "$STAGE"
find . -name '*.cpp' ! -path './Main/*' ! -path './include/*' | sed 's#^\./##' > srcs.txt
mkdir -p obj
cat srcs.txt | xargs -P "$JOBS" -I{} sh -c 'o=obj/$(echo {} | tr "/" "_").o; g++ '"$CXXFLAGS"' -c {} -o $o || exit 255'
# harness (ours) - compiled against the staged reference headers and our C-ABI header
PKG="$HERE/../pysicalbasedraytracer_b200"
g++ $CXXFLAGS -I"$HERE/../include" -I"$PKG/host" -c "$HERE/ref_harness.cpp" -o obj/ref_harness.o
# libpbr_ref.so = the reference + the harness: NO product code (the reference arm of bench.py loads only this)
g++ -shared -fopenmp -o "$OUT/libpbr_ref.so" obj/*.o
echo "built $OUT/libpbr_ref.so"
# libdropin_test.so = the C++ drop-in test: the ONLY oracle-side object that links the product library
if [ -f "$PKG/libpbrgpu.so" ]; then
  g++ $CXXFLAGS -I"$HERE/../include" -I"$PKG/host" -c "$HERE/dropin_test.cpp" -o dropin_test.o
  g++ -shared -fopenmp -o "$OUT/libdropin_test.so" dropin_test.o -L"$OUT" -lpbr_ref -L"$PKG" -lpbrgpu \
      -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../pysicalbasedraytracer_b200'
  echo "built $OUT/libdropin_test.so"
fi
