#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own tools / geometry / signal_processing sources, where they lie
# under /root/reference, into oracle/_ref/libs/*.so with the flags the reference's CMake
# effectively uses (Release: -O3 -DNDEBUG, C++14, no -fopenmp, no -march).
# Nothing is copied out of /root/reference; only the built .so files land in oracle/_ref/
# (git-ignored, shipped to the GPU box by gpurun). The product never links or loads them.
set -euo pipefail
R="${LIBRIR_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$R/src/cpp/signal_processing" ]; then
  echo "build_ref: $R not present; keeping prebuilt oracle/_ref if any" >&2
  exit 0
fi
mkdir -p "$OUT/libs" "$OUT/cfg"
sed -e 's/@PROJECT_NAME@/librir/' -e 's/@PROJECT_VERSION@/6.1.2/' \
    -e 's/@PROJECT_VERSION_MAJOR@/6/' -e 's/@PROJECT_VERSION_MINOR@/1/' \
    -e 's/@PROJECT_VERSION_PATCH@/2/' "$R/rir_config.h.in" > "$OUT/cfg/rir_config.h"
CXX="${ORACLE_CXX:-/usr/bin/g++}"   # not $CXX: the image's /opt/gcc wrapper lacks libgomp.spec
FLAGS="-std=c++14 -O3 -DNDEBUG -fPIC -shared -w"
INC="-I$OUT/cfg -I$HERE/shim -I$R/src/cpp/tools -I$R/src/cpp/geometry -I$R/src/cpp/signal_processing"
build() { # name define sources... -- extra link flags
  local name="$1" def="$2"; shift 2
  if [ "$OUT/libs/lib$name.so" -nt "$0" ] && [ -z "${FORCE:-}" ]; then return; fi
  echo "build_ref: lib$name.so"
  $CXX $FLAGS -D$def $INC "$@" -o "$OUT/libs/lib$name.so"
}
build tools BUILD_TOOLS_LIB -DHAVE_ZSTD -DZSTD_COMPRESS $R/src/cpp/tools/*.cpp -l:libzstd.so.1 -lpthread
build geometry BUILD_GEOMETRY_LIB $R/src/cpp/geometry/*.cpp -L"$OUT/libs" -ltools -Wl,-rpath,'$ORIGIN'
build signal_processing BUILD_SIGNAL_PROCESSING_LIB $R/src/cpp/signal_processing/*.cpp -L"$OUT/libs" -ltools -Wl,-rpath,'$ORIGIN'
# all-cores variant of the same sources (reported baseline only, never the parity oracle)
if [ ! "$OUT/libs/libsignal_processing_omp.so" -nt "$0" ] || [ -n "${FORCE:-}" ]; then
  echo "build_ref: libsignal_processing_omp.so"
  $CXX $FLAGS -fopenmp -DBUILD_SIGNAL_PROCESSING_LIB $INC $R/src/cpp/signal_processing/*.cpp \
      -L"$OUT/libs" -ltools -Wl,-rpath,'$ORIGIN' -o "$OUT/libs/libsignal_processing_omp.so"
fi
if [ ! "$OUT/libs/libref_shim.so" -nt "$HERE/ref_shim.cpp" ] || [ -n "${FORCE:-}" ]; then
  echo "build_ref: libref_shim.so"
  $CXX $FLAGS $INC "$HERE/ref_shim.cpp" -L"$OUT/libs" -lsignal_processing -ltools -Wl,-rpath,'$ORIGIN' -o "$OUT/libs/libref_shim.so"
fi
# the zstd movie file (ZFile.cpp is the one video_io source that needs nothing but the tools library)
if [ ! "$OUT/libs/libref_zfile.so" -nt "$HERE/zfile_shim.cpp" ] || [ ! "$OUT/libs/libref_zfile.so" -nt "$0" ] || [ -n "${FORCE:-}" ]; then
  echo "build_ref: libref_zfile.so"
  $CXX $FLAGS -DBUILD_IO_LIB $INC -I$R/src/cpp/video_io "$R/src/cpp/video_io/ZFile.cpp" "$HERE/zfile_shim.cpp" \
      -L"$OUT/libs" -ltools -Wl,-rpath,'$ORIGIN' -o "$OUT/libs/libref_zfile.so"
fi
echo "build_ref: done -> $OUT/libs"
