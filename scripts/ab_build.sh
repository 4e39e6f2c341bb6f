#!/usr/bin/env bash
# Builds alternative libraries for A/B timing (never shipped): natrix_b200/_ab/lib_<name>.so with extra nvcc flags.
# usage: bash scripts/ab_build.sh name "-DNATRIX_TB_FMA=0" [name2 "flags2" ...];  run with NATRIX_B200_LIB=<path>
set -e
cd "$(dirname "$0")/../natrix_b200/csrc"
mkdir -p ../_ab
while [ $# -ge 2 ]; do
  make -s BUILD=build_ab_$1 OUT=../_ab/lib_$1.so EXTRA="$2" > /dev/null
  echo "built natrix_b200/_ab/lib_$1.so ($2)"
  shift 2
done
