#!/bin/bash
# Dev tool: compute-sanitizer (memcheck, racecheck, synccheck, initcheck) over tools/sanitize_driver.py.
out=${1:-gpurun_out/sanitizer.txt}
: > $out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== $tool" >> $out
  timeout 1200 compute-sanitizer --tool $tool python tools/sanitize_driver.py 2>&1 | grep -E "unwrapped energy|gradient|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Invalid|Uninit" | head -40 >> $out
done
cat $out
