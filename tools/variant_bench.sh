#!/bin/bash
# usage: tools/variant_bench.sh <lib.so> [env assignments...]  -- one short bench line per variant (tuning aid)
lib=$1; shift
env THINCURR_B200_LIB=$lib "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('%-40s ms/step %.2f  pairs/s %.3e  frac %.3f  launches %d' % ('$lib $*', d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches']))"
