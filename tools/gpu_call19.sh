#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python tools/cfg2_after_big.py 2>&1 | tail -8
