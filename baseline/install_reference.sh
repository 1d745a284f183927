#!/bin/bash
# Installs the UNMODIFIED reference (danieleschmidt/Graph-Hypernetwork-Forge, pure Python on torch) under
# baseline/_ref so that `bench.py --impl reference` can time it on the GPU box's host cores.  baseline/_ref is
# git-ignored (no reference source enters the history) but travels with the gpurun snapshot.  The reference tree is
# read-only, so pip builds the wheel from a copy under /tmp; --no-deps: its only dependency, torch, is in the image.
set -e
cd "$(dirname "$0")/.."
REF=${1:-/root/reference}
[ -d "$REF" ] || { echo "no reference tree at $REF: keeping what baseline/_ref holds"; exit 0; }
rm -rf /tmp/ghf_refcopy baseline/_ref
cp -r "$REF" /tmp/ghf_refcopy
python -m pip install --quiet --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target baseline/_ref /tmp/ghf_refcopy
rm -rf /tmp/ghf_refcopy
ls baseline/_ref
