#!/bin/bash
# Dev tool: bench.py over column-pass variants / precisions / meshes (one JSON line each -> gpurun_out/variants.jsonl)
out=${1:-gpurun_out/variants.jsonl}
: > "$out"
run() { echo "# $*" >> "$out"; timeout 120 python bench.py --no-cpu --steps 100 --warmup 10 "$@" 2>>"$out.err" | python tools/bench_summ2.py >> "$out"; }
for tile in 0 3; do
  run --col-tile $tile
  run --col-tile $tile --precision c64
  run --col-tile $tile --mode real
  run --col-tile $tile --mesh 1024
  run --col-tile $tile --mesh 4096 --steps 40
done
