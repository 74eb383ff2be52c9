#!/bin/sh
# oracle/_ref: the UNMODIFIED reference sources the CPU arm and the live-harness test execute on the GPU box.
#
# The reference is pure Python (no build system, nothing to compile): the "build" of oracle/_ref is a verbatim copy of
# the files on the hot path and of its consumers, taken from where they lie under /root/reference.  oracle/_ref/ is
# git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so it travels to the GPU box like the
# built .so files.  __graft_entry__.build() runs this whenever /root/reference is present.
#
#   config.py                  ConfigEuRoC                                   (src/config.py)
#   image_processing/          ImageProcessor.stereo_callback, the hot path  (src/image_processing/pipeline.py:46-150)
#   msckf.py feature/ utils.py the consumer of feature_msg                   (src/msckf.py:177-228)
#   modules/vio.py             the thread harness                            (src/modules/vio.py:6-53)
#   streaming/                 EuRoC reader, paced publisher                 (src/streaming/dataset.py, publisher.py)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REFERENCE_SRC:-/root/reference/src}"
DST="$ROOT/oracle/_ref"
if [ ! -d "$REF" ]; then
    echo "make_oracle_ref: $REF not present (GPU box): keeping $DST as it travelled" >&2
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/modules"
cp "$REF/config.py" "$REF/msckf.py" "$REF/utils.py" "$DST/"
cp -r "$REF/image_processing" "$REF/feature" "$REF/streaming" "$DST/"
cp "$REF/modules/vio.py" "$DST/modules/"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
# provenance: a digest of every copied file, checked by tests/test_oracle_ref.py against the tree it was copied from
(cd "$DST" && find . -type f -name '*.py' | sort | xargs sha256sum) > "$DST/SHA256SUMS"
echo "oracle/_ref: $(grep -c . "$DST/SHA256SUMS") files copied unmodified from $REF"
