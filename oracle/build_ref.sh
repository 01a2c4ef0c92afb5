#!/bin/bash
# TEST / BENCH INFRASTRUCTURE.  Stages the UNMODIFIED reference (TUM-DAML/gemnet_pytorch, pure Python: nothing to
# compile) from /root/reference into the git-ignored oracle/_ref/ so that it travels to the GPU box with the tree:
#   oracle/_ref/gemnet/**            the reference package, byte for byte
#   oracle/_ref/pretrained/**        shipped checkpoints + model kwargs + scaling factors (GemNet-T / Q / dT / dQ)
#   oracle/_ref/scaling_factors.json
# Used ONLY by tests/ (acceptance tests: the reference's own Trainer on top of gemnet_pytorch_b200.GemNet) and by
# `bench.py --impl reference` / the cpu_baseline leg (kind "reference": the reference's own CPU forward through
# oracle/ref_shims.py, four import shims that touch no hot-path arithmetic).  Nothing under gemnet_pytorch_b200/ reads it.
# Never committed (.gitignore: oracle/_ref/); a no-op when /root/reference is absent (the GPU box uses the staged copy).
set -e
SRC=${GEMNET_REFERENCE_SRC:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$SRC/gemnet/model" ]; then
  echo "build_ref: $SRC not present; keeping $DST as is" >&2
  exit 0
fi
mkdir -p "$DST"
rm -rf "$DST/gemnet" "$DST/pretrained"
cp -r "$SRC/gemnet" "$DST/gemnet"
cp -r "$SRC/pretrained" "$DST/pretrained"
cp "$SRC/scaling_factors.json" "$DST/scaling_factors.json" 2>/dev/null || true
find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} + 2>/dev/null || true
( cd "$SRC" && find gemnet pretrained -type f | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "build_ref: staged $(find "$DST/gemnet" -name '*.py' | wc -l) reference modules into $DST"
