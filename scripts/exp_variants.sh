#!/bin/bash
# run a command once per experiment build of the library (librir_b200/libs_exp/<name>/), swapping it into place
# usage: scripts/exp_variants.sh "<name> <name> ..." <command...>   (the default build is restored afterwards)
set -u
names="$1"; shift
L=librir_b200/libs/libsignal_processing_b200.so
cp -p $L /tmp/default_lib.so
for n in default $names; do
  if [ "$n" = default ]; then cp -p /tmp/default_lib.so $L; else cp -p librir_b200/libs_exp/$n/libsignal_processing_b200.so $L; fi
  echo "### $n"
  "$@"
done
cp -p /tmp/default_lib.so $L
