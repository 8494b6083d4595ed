#!/bin/bash
# (inside gpurun) the same_size leg of bench.py at a reduced main size, with and without the axpby batching
for V in "" "B200_NO_AXPBY_BATCH=1"; do
	env $V timeout 400 python bench.py --m ${1:-100} --steps 1 --warmup 1 --no-parity 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['same_size']
print('$V', d['value'], 'same_size b200_s', s['b200_s'], 'e2e', s['b200_e2e_s'], 'ref', s['reference_s'])
"
done
