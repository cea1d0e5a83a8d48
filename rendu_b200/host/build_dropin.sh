#!/usr/bin/env bash
# Builds build/pathtracer_b200: the reference's own engine objects (scene loader, resources, lights — compiled from
# /root/reference by oracle/build_ref.sh) linked with the GPU Raycaster/PathTracer of this directory INSTEAD of the
# reference's Raycaster.cpp / PathTracer.cpp. This is what swapping the two classes inside Rendu amounts to.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REF:-/root/reference}"
OBJ="$ROOT/oracle/_ref/obj"
if [ ! -d "$REF/src/engine" ] || [ ! -d "$OBJ" ]; then
	echo "build_dropin: needs $REF and oracle/_ref (run oracle/build_ref.sh); keeping any prebuilt binary" >&2
	exit 0
fi
mkdir -p "$ROOT/build/dropin"
CXX=/usr/bin/g++
# These directories go FIRST so that "raycaster/Raycaster.hpp" and "PathTracer.hpp" resolve to the drop-in headers.
INC="-I $HERE/shadow -I $HERE -I $ROOT/include -I $ROOT/oracle/ref_harness/stub -I $REF/src/engine -I $REF/src/libs -I $REF/src/libs/glfw/include"
for f in Raycaster PathTracer dropin_main; do
	$CXX -std=c++11 -O2 -w $INC -c "$HERE/$f.cpp" -o "$ROOT/build/dropin/$f.o"
done
# The reference's BVH debug view, UNMODIFIED, compiled against the drop-in header: it is a friend of Raycaster and reads
# _hierarchy / _triangles / _vertices directly (RaycasterVisualisation.cpp:15-30,41-77), which is why the mirror exists.
$CXX -std=c++11 -O2 -w $INC -c "$REF/src/engine/raycaster/RaycasterVisualisation.cpp" -o "$ROOT/build/dropin/RaycasterVisualisation.o"
REFOBJS=$(ls "$OBJ"/*.o | grep -v -E "/(Raycaster|PathTracer|ref_api|ref_build_scene)\.o$")
$CXX -o "$ROOT/build/pathtracer_b200" "$ROOT/build/dropin/"{Raycaster,PathTracer,dropin_main,RaycasterVisualisation}.o $REFOBJS \
	-L"$ROOT/rendu_b200" -lrendu_b200 -Wl,-rpath,'$ORIGIN/../rendu_b200' -lpthread
echo "build_dropin: $ROOT/build/pathtracer_b200 ready"
