#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the real reference (kosua20/Rendu) CPU path tracer as a checker.
#
# Compiles the reference's own, unmodified translation units *where they lie* under $REF (default
# /root/reference) into oracle/_ref/ (git-ignored; it travels to the GPU box with the snapshot) and links them
# with the harness in oracle/ref_harness/ into oracle/_ref/librendu_ref.so. No reference source is copied into
# the repository; the only text sliced out of the reference tree is the body of Texture::sampleCubemap
# (src/engine/resources/Texture.cpp:141-191), extracted at build time into oracle/_ref/gen/ because the rest of
# that file needs Vulkan. The reference's premake/Vulkan build system is not run.
#
# Flags follow the reference's release configuration (premake5.lua:49-51,66-69: C++11, optimize On, no -march,
# no fast-math) plus -ffp-contract=off so that no host ever fuses multiply-adds.
#
# The bundled scenes/meshes/textures the reference loads at run time (resources/common, resources/pbrdemo) are
# mirrored into oracle/_ref/resources/ so that the reference can also be timed on the GPU box, which has no
# /root/reference.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
SRC="$REF/src"
if [ ! -d "$SRC/engine/raycaster" ]; then
	echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)" >&2
	exit 0
fi
mkdir -p "$OUT/obj" "$OUT/gen"

# The image exports CXX=/opt/gcc/bin/g++, a wrapper that links libstdc++ statically; a second libstdc++ inside a
# python process breaks iostreams, so the system compiler is named explicitly.
CXX="${REF_CXX:-/usr/bin/g++}"
CC="${REF_CC:-/usr/bin/gcc}"
CXXFLAGS="-std=c++11 -O2 -ffp-contract=off -fPIC -w"
INC="-I $HERE/ref_harness/stub -I $OUT/gen -I $SRC/engine -I $SRC/libs -I $SRC/libs/glfw/include -I $SRC/apps/pathtracer"

UNITS="
engine/raycaster/Raycaster engine/raycaster/Intersection
engine/resources/Bounds engine/resources/Image engine/resources/Mesh engine/resources/ResourcesManager engine/resources/Library
engine/generation/Random
engine/system/Logger engine/system/Query engine/system/TextUtilities engine/system/Config engine/system/Codable
engine/input/Camera
engine/scene/Scene engine/scene/Object engine/scene/Material engine/scene/Sky engine/scene/Animation engine/scene/LightProbe
engine/scene/lights/Light engine/scene/lights/PointLight engine/scene/lights/SpotLight engine/scene/lights/DirectionalLight
apps/pathtracer/PathTracer apps/pathtracer/MaterialGGX apps/pathtracer/MaterialSky
"

# Body of Texture::sampleCubemap, sliced verbatim from the read-only tree (never committed).
awk '/^glm::vec3 Texture::sampleCubemap/{on=1} on{print} on&&/^}/{exit}' "$SRC/engine/resources/Texture.cpp" > "$OUT/gen/sampleCubemap.inc"
grep -q "sampleCubemap" "$OUT/gen/sampleCubemap.inc"

pids=()
compile() { # src obj compiler flags
	if [ ! -f "$2" ] || [ "$1" -nt "$2" ]; then
		$3 $4 $INC -c "$1" -o "$2" &
		pids+=($!)
	fi
}
for u in $UNITS; do
	compile "$SRC/$u.cpp" "$OUT/obj/$(basename "$u").o" "$CXX" "$CXXFLAGS"
done
compile "$SRC/libs/miniz/miniz.c" "$OUT/obj/miniz.o" "$CC" "-O2 -fPIC -w"
compile "$SRC/libs/mikktspace/mikktspace.c" "$OUT/obj/mikktspace.o" "$CC" "-O2 -fPIC -w"
compile "$HERE/ref_harness/ref_api.cpp" "$OUT/obj/ref_api.o" "$CXX" "$CXXFLAGS"
compile "$HERE/ref_harness/ref_shim.cpp" "$OUT/obj/ref_shim.o" "$CXX" "$CXXFLAGS"
compile "$HERE/ref_harness/ref_build_scene.cpp" "$OUT/obj/ref_build_scene.o" "$CXX" "$CXXFLAGS"
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done

WRAP="_ZNK9Raycaster10intersectsERKN3glm3vecILi3EfLNS0_9qualifierE0EEES5_ff"
$CXX -shared -o "$OUT/librendu_ref.so" "$OUT"/obj/*.o -Wl,--wrap=$WRAP -Wl,--wrap=_ZNSt6thread20hardware_concurrencyEv -Wl,--no-undefined -Wl,-Bsymbolic -lpthread

# Second checker library: the reference's own RaycasterVisualisation (unmodified) + its harness. Separate because ref_api.cpp
# defines a stand-in class of that name (the friend hook); it shares every other reference object.
mkdir -p "$OUT/objvis"
pids=()
compile "$SRC/engine/raycaster/RaycasterVisualisation.cpp" "$OUT/objvis/RaycasterVisualisation.o" "$CXX" "$CXXFLAGS"
compile "$HERE/ref_harness/refvis_api.cpp" "$OUT/objvis/refvis_api.o" "$CXX" "$CXXFLAGS"
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
VISOBJS=$(ls "$OUT"/obj/*.o | grep -v "/ref_api.o$")
$CXX -shared -o "$OUT/librendu_refvis.so" $VISOBJS "$OUT"/objvis/*.o -Wl,--no-undefined -Wl,-Bsymbolic -lpthread

# Run-time data of the reference (not source code): scenes, meshes, textures.
mkdir -p "$OUT/resources"
for d in common pbrdemo; do
	if [ ! -d "$OUT/resources/$d" ]; then
		mkdir -p "$OUT/resources/$d"
		( cd "$REF/resources/$d" && find . -type f ! -path "*/shaders/*" -print0 | cpio -0pdm --quiet "$OUT/resources/$d" ) 2>/dev/null \
			|| cp -r "$REF/resources/$d/." "$OUT/resources/$d/"
	fi
done
echo "build_ref: $OUT/librendu_ref.so ready"
