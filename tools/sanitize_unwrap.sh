#!/bin/bash
# Dev tool: compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over tools/sanitize_unwrap.py.
out=${1:-gpurun_out/sanitizer_unwrap.txt}
: > $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" >> $out
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_unwrap.py 2>&1 | grep -E "same field|DIFFERENT|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Invalid|Uninit" | head -40 >> $out
done
cat $out
