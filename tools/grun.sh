#!/bin/bash
# retry wrapper around gpurun: exit code 3 (no box / slot free) is retried every 90 s, up to 20 times
# usage: tools/grun.sh [--gpus N] <timeout-seconds> '<command>'
extra=()
if [ "$1" == "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "${extra[@]}" --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
