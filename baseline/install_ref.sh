#!/bin/sh
# Install the UNMODIFIED reference arithmetic path into baseline/_ref (git-ignored; travels to the GPU box with the repo
# snapshot): the vendored OpenCLIP 2.24.0 of lmb-freiburg/understanding-clip-ood (deps/open_clip: packages `open_clip` and
# `training`).  Build container only -- /root/reference does not exist on the GPU box.
#
# The top-level `xclip` package of the reference cannot be installed offline (its pyproject.toml needs the `hatchling` build
# backend, which is not in the image, and pins torch 2.4.1 / lightning / textacy ...): the three lines of
# xclip/zero_shot.py that sit on the hot path (F.normalize, tensordot, argmax / topk: :42-60, :103-109) are issued by
# bench.py's reference arm on top of the reference model instead.
set -e
REF=${REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -d "$REF/deps/open_clip" ] || { echo "no reference at $REF: nothing to install"; exit 0; }
TMP=$(mktemp -d)
cp -r "$REF/deps/open_clip" "$TMP/open_clip_src"          # the build writes into the source tree; /root/reference is read-only
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP/open_clip_src"
rm -rf "$TMP"
ls "$HERE/_ref"
