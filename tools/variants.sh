#!/bin/bash
# Times kernel-variant builds (build/librcvvote_*.so) against the default library.
for v in default "$@"; do
  if [ "$v" = default ]; then unset RCV_LIB_PATH; else export RCV_LIB_PATH=$PWD/build/librcvvote_$v.so; fi
  echo "== variant: $v"
  timeout 300 python bench.py --no-e2e --no-cpu --steps 3 > /tmp/b.json 2>/tmp/b.err || tail -3 /tmp/b.err
  python -c "
import json
d=json.load(open('/tmp/b.json'))
print('frames/s %.0f  Gvotes/s %.1f  k_vote ms %.1f  step ms %.1f  frac %.4f' % (d['value'], d['gvotes_per_s'], d['roofline']['kernel_ms_per_launch'], d['ms_per_step'], d['roofline']['frac']))"
done
