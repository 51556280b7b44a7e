#!/bin/sh
# TEST / BENCH INFRASTRUCTURE ONLY.  Places a verbatim copy of the UNMODIFIED reference package
# (pure Python, no build step) under oracle/_ref/ so that it travels to the GPU box, where
# /root/reference does not exist.  oracle/_ref/ is git-ignored: reference sources never enter
# this repository's history.  Used by `bench.py --impl reference` / `cpu_baseline` (kind
# "reference") and by tests/test_dropin_gpu.py (stock mici samplers over mici_b200 integrators).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
src="${MICI_REFERENCE_SRC:-/root/reference/src/mici}"
if [ ! -d "$src" ]; then
  echo "build_ref: $src not present (GPU box?) -- keeping existing oracle/_ref" >&2
  exit 0
fi
rm -rf "$here/_ref"
mkdir -p "$here/_ref"
cp -r "$src" "$here/_ref/mici"
find "$here/_ref" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
( cd "$src/../.." && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$here/_ref/COMMIT"
echo "build_ref: copied $src -> $here/_ref/mici ($(cat "$here/_ref/COMMIT"))"
