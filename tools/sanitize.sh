#!/bin/bash
# compute-sanitizer over the encode / decode / trainer paths (race detection / sanitizers).
#   tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck ...]      (default: memcheck racecheck initcheck)
# Each tool runs tools/sanitize_driver.py: small inputs of every shape that reaches a different kernel (short and long
# pre-tokens, added tokens inside words, NFC, documents that start inside a slice, cache publication under contention,
# decode with clean-up, trainer).  Logs: gpurun_out/sanitize_<tool>.log ; the summary line of each goes to stdout.
# The things to prove: the lock-free publication of cache slots (CAS -> plain stores -> __threadfence -> volatile store,
# readers use ld.global.cg: encode_fused.cu) and the trainer's shared-table CAS loops.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TOOLS="${*:-memcheck racecheck initcheck}"
rc=0
for t in $TOOLS; do
    extra=""
    [ "$t" = initcheck ] && extra="--track-unused-memory no"
    timeout 900 compute-sanitizer --tool "$t" $extra --error-exitcode 9 --print-limit 20 python tools/sanitize_driver.py > "gpurun_out/sanitize_$t.log" 2>&1
    r=$?
    [ $r -ne 0 ] && rc=$r
    echo "== $t: exit $r :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "gpurun_out/sanitize_$t.log" | tail -1)"
    grep -E "driver:" "gpurun_out/sanitize_$t.log" | tail -3
done
exit $rc
