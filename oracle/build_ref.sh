#!/usr/bin/env bash
# oracle/build_ref.sh — stage the UNMODIFIED reference (pure Python) where the GPU box can import it.
#
# The reference (funcwj/aps) is PyTorch-on-CPU code; `/root/reference` exists only in the build container.  This
# recipe copies its `aps/` package byte for byte into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it travels to
# the GPU box like the built .so) next to the three stand-in third-party modules of oracle/ref_shims/ (librosa.filters,
# kaldi_python_io, soundfile — not installed in this image).  bench.py's `--impl reference` arm and
# its `cpu_baseline` leg then time the stock `aps.transform.AsrTransform` / `aps.asr.transformer.encoder.
# TransformerEncoder` / ... on the box's host cores (cpu_baseline.kind = "reference").  Nothing under oracle/_ref is
# ever imported by the product (tests/test_layout.py).  No reference source enters the git history.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${APS_REFERENCE:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/aps" ]; then
    echo "build_ref: $SRC/aps not found (GPU box: the staged copy in oracle/_ref is used as it is)"
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/aps" "$DST/aps"
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
for shim in librosa kaldi_python_io soundfile; do
    cp -r "$HERE/ref_shims/$shim" "$DST/$shim"
done
( cd "$SRC" && { git rev-parse HEAD 2>/dev/null || echo "no-git"; } ) > "$DST/REFERENCE_REV"
( cd "$DST/aps" && find . -name "*.py" -print0 | sort -z | xargs -0 sha1sum | sha1sum | cut -d' ' -f1 ) > "$DST/SOURCE_SHA1"
echo "build_ref: staged $(find "$DST/aps" -name '*.py' | wc -l) reference files into $DST (tree sha1 $(cat "$DST/SOURCE_SHA1"))"
