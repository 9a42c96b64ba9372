#!/usr/bin/env bash
# Stage the UNMODIFIED reference under baseline/_ref/ (git-ignored, NOT gpurun-ignored) so that it travels to the GPU
# box with the snapshot.  Used only by measurement tools (tools/reference_gpu.py: the reference's own PyTorch/CUDA op
# path as GPU baseline; tools/train_step.py: the reference's step functions from its own source).
# Nothing under 3d-fm-gan_b200/, tests -m gpu's required set, smoke() or bench.py depends on it.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
DST="$ROOT/baseline/_ref"
[ -f "$SRC/stylegan2.py" ] || { echo "no reference checkout at $SRC" >&2; exit 1; }
# keep a pre-built extension directory (tools/reference_gpu.py jit) across re-staging
KEEP=""
if [ -d "$DST/_torch_ext" ]; then KEEP="$(mktemp -d)"; mv "$DST/_torch_ext" "$KEEP/"; fi
rm -rf "$DST"
mkdir -p "$DST"
if [ -n "$KEEP" ]; then mv "$KEEP/_torch_ext" "$DST/"; rmdir "$KEEP"; fi
# code only: python sources, the two op extensions, the small LPIPS linear-layer weights; no docs / env files
( cd "$SRC" && tar cf - --exclude='doc' --exclude='Conda_Env_Setup' --exclude='DiscoFaceGAN_related_scripts' \
      --exclude='.git' --exclude='__pycache__' . ) | ( cd "$DST" && tar xf - )
echo "staged $(find "$DST" -type f | wc -l) files ($(du -sh "$DST" | cut -f1)) from $SRC into baseline/_ref"
