#!/bin/bash
# Quick GPU check: parity tests then a short bench.  Usage: tools/gpu_quick.sh <tag> [bench args...]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
timeout 900 python bench.py --steps 3 --no-cpu --e2e-steps 1 "$@" 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('frames/s %.0f  Gvotes/s %.1f  k_vote ms %.1f  step ms %.1f  frac %.4f  e2e %s' % (d['value'], d['gvotes_per_s'], d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['roofline']['frac'], d['e2e'] and d['e2e']['value']))"
tail -3 gpurun_out/${TAG}_bench.err
