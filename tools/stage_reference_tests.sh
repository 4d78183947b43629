#!/bin/bash
# Build container only: stage the reference's OWN test files into the git-ignored scratch directory
# baseline/_ref/tests (it travels to the GPU box with the gpurun snapshot; nothing is copied into the
# repository history). The tests import `hpc` from ../build/lib.*/ relative to their directory, so a
# link to this repo's package is placed there.
set -e
cd "$(dirname "$0")/.."
mkdir -p baseline/_ref/build/lib.b200
rm -rf baseline/_ref/tests
cp -r /root/reference/tests baseline/_ref/tests
ln -sfn ../../../../hpc-ops_b200/hpc baseline/_ref/build/lib.b200/hpc
echo "staged $(ls baseline/_ref/tests | wc -l) files"
