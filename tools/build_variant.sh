#!/bin/bash
# builds a compile-time variant of libb2pbr.so into build_variants/<name>.so (probes load it through PROBE_LIB);
# usage: tools/build_variant.sh name VAR=value ...
set -e
name=$1; shift
cp dplug_b200/libb2pbr.so /tmp/libb2pbr.keep
env "$@" python dplug_b200/build.py --force > /dev/null
cp dplug_b200/libb2pbr.so build_variants/$name.so
python dplug_b200/build.py --force > /dev/null
cmp -s dplug_b200/libb2pbr.so /tmp/libb2pbr.keep || echo "note: default build differs from the one before (sources changed since)"
