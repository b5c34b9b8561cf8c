#!/bin/bash
# Development helper: rebuild the product library and the instrumented variant from the repo root.
set -e
cd "$(dirname "$0")/.."
python -c "
from clibd_b200 import _build
print(_build.build(force=True))
print(_build.build_variant('timing', ['CLIBD_BWD_TIMING']))"
