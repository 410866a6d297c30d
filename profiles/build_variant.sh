#!/bin/bash
# usage: profiles/build_variant.sh KE MINB KV [TR]  -> build_variants/lib{TR|NT}_KE_MINB_KV.so (diagnostic builds of the EV kernel)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TAG=${4:-NT}
EX="-DSG_EV_KE=$1 -DSG_EV_MINB=$2 -DSG_EV_KV=$3 ${SG_EXTRA_DEFS:-}"
if [ "$TAG" = TR ]; then EX="$EX -DSG_EV_TRACE"; fi
mkdir -p "$ROOT/build_variants"
SG_LIB_PATH="$ROOT/build_variants/lib${TAG}_$1_$2_$3${SG_SUFFIX:-}.so" SG_NVCC_EXTRA="$EX" python - <<PY
import sys
sys.path.insert(0, "$ROOT")
from sustaingym_b200 import _build
_build.BUILD_DIR = "$ROOT/build_variants/b${TAG}_$1_$2_$3${SG_SUFFIX:-}"
print(_build.build(force=True))
PY
