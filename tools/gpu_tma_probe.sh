#!/bin/bash
cd "$(dirname "$0")/.."
for v in 0 1 2 3 4 5; do timeout 60 ./tools/experimental/tma_probe $v 2>&1 | tr '\n' ' '; echo; done
