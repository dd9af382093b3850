#!/bin/sh
# Unpacks the reference's own prebuilt binary (pysplashsurf 0.14.0 manylinux wheel, shipped inside the
# reference tree) into oracle/_ref/.  The reference is Rust and there is no Rust toolchain in this image,
# so the wheel -- built by the reference's authors from the same v0.14.0 sources -- is the runnable
# reference.  oracle/_ref/ is git-ignored (not product source) but travels to the GPU box with gpurun.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
WHL=/root/reference/splashsurf_studio/src/wheels/pysplashsurf-0.14.0.0-cp310-abi3-manylinux_2_17_x86_64.manylinux2014_x86_64.whl
if [ -f "$HERE/_ref/pysplashsurf/pysplashsurf.abi3.so" ]; then exit 0; fi
if [ ! -f "$WHL" ]; then echo "reference wheel not present; oracle/_ref not built" >&2; exit 0; fi
mkdir -p "$HERE/_ref"
python3 -m zipfile -e "$WHL" "$HERE/_ref"
