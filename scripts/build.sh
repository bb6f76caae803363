#!/bin/bash
# Incremental in-tree build of the CUDA library (run from anywhere).  scripts/build.sh [-v] [--force]
cd "$(dirname "$0")/../gym-formation_b200" && exec python -m formation_gym._build "$@"
