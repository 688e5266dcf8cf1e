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
# video_io: the reference's writer / reader sources against its own ffmpeg 7.1 headers, with the 46 libav entry points
# they call provided by oracle/libav_stub.c (identity codec + trivial container; ffmpeg / x264 / kvazaar themselves are
# fetched from the network by the reference's build and are not available here).  Objects are built in parallel.
VIO_SRC="BaseCalibration IRFileLoader h264 IRVideoLoader HCCLoader video_io ZFile"
FF="$R/extra/ffmpeg/ffmpeg-7.1-msvc/include"
if [ ! "$OUT/libs/libvideo_io.so" -nt "$HERE/libav_stub.c" ] || [ ! "$OUT/libs/libvideo_io.so" -nt "$0" ] || [ -n "${FORCE:-}" ]; then
  echo "build_ref: libvideo_io.so (stub libav)"
  mkdir -p "$OUT/obj"
  pids=""
  for f in $VIO_SRC; do
    $CXX -std=c++14 -O3 -DNDEBUG -fPIC -w -DBUILD_IO_LIB -DUSE_ZFILE $INC -I$R/src/cpp/video_io -I"$FF" \
        -c "$R/src/cpp/video_io/$f.cpp" -o "$OUT/obj/$f.o" & pids="$pids $!"
  done
  ${ORACLE_CC:-/usr/bin/gcc} -std=gnu11 -O2 -fPIC -w -I"$FF" -c "$HERE/libav_stub.c" -o "$OUT/obj/libav_stub.o" & pids="$pids $!"
  for p in $pids; do wait $p; done
  objs=""; for f in $VIO_SRC libav_stub; do objs="$objs $OUT/obj/$f.o"; done
  $CXX -shared -o "$OUT/libs/libvideo_io.so" $objs -L"$OUT/libs" -lsignal_processing -lgeometry -ltools \
      -Wl,-rpath,'$ORIGIN' -Wl,--no-undefined
  rm -rf "$OUT/obj"
fi
# the reference's own Python package, laid out the way its wheel is (librir/ + librir/libs/*.so), so that the GPU box
# (where /root/reference does not exist) can run the UNMODIFIED wrappers: once on the compiled reference
# (oracle/_ref/pkg/librir) and once with this repo's libraries swapped in (tests/test_dropin_*.py copy the tree and
# replace libs/).  oracle/_ref is git-ignored: nothing of the reference enters the history.
if [ ! -f "$OUT/pkg/librir/__init__.py" ] || [ "$OUT/libs/libvideo_io.so" -nt "$OUT/pkg/librir/libs/libvideo_io.so" ] || [ -n "${FORCE:-}" ]; then
  echo "build_ref: python package -> $OUT/pkg/librir"
  rm -rf "$OUT/pkg"
  mkdir -p "$OUT/pkg"
  cp -r "$R/src/python/librir" "$OUT/pkg/librir"
  mkdir -p "$OUT/pkg/librir/libs"
  for l in tools geometry signal_processing video_io; do cp "$OUT/libs/lib$l.so" "$OUT/pkg/librir/libs/"; done
fi
# the reference's own Python tests (tests/python), laid out as the package `tests.python` they import themselves as, so that
# the GPU box can run them UNMODIFIED on top of the drop-in libraries (tests/test_reference_own_tests.py).  Git-ignored like
# the rest of oracle/_ref; the 1 MB geometry fixture (circle.py) is left out.
if [ ! -f "$OUT/reftests/tests/python/conftest.py" ] || [ ! -f "$OUT/reftests/examples/ir_saver.py" ] || [ -n "${FORCE:-}" ]; then
  echo "build_ref: reference tests -> $OUT/reftests"
  rm -rf "$OUT/reftests"
  mkdir -p "$OUT/reftests/tests/python"
  : > "$OUT/reftests/tests/__init__.py"
  for f in __init__.py conftest.py test_IRMovie.py test_rir.py test_video_io.py test_registration.py test_FileAttributes.py; do
    cp "$R/tests/python/$f" "$OUT/reftests/tests/python/"
  done
  mkdir -p "$OUT/reftests/examples"   # the reference's example scripts of the path (saver / reader, registration)
  cp "$R/examples/ir_saver.py" "$R/examples/registration.py" "$OUT/reftests/examples/"
fi
echo "build_ref: done -> $OUT/libs"
