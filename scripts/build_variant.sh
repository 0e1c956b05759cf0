#!/bin/bash
# scripts/build_variant.sh NAME [extra nvcc flags...]  -> quantumattention_b200/libqattn_sm100_NAME.so  (dev A/B builds)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
python - "$name" "$@" <<'PY'
import sys
from quantumattention_b200 import build as b
import os
name, extra = sys.argv[1], sys.argv[2:]
out = os.path.join(b.HERE, f"libqattn_sm100_{name}.so")
b.compile_and_link(out, [os.path.join(b.CSRC, s) for s in b.SOURCES], extra=extra)
print("built", out)
PY
