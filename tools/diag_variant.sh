#!/bin/bash
# usage: tools/diag_variant.sh <lib.so> : run bench with a prebuilt library variant
cp "$1" complexity-tokenizer_b200/complexity_tokenizer/libctk.so
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'ms', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['all_kernels_ms_per_step'].items() if v>0.1})
"
