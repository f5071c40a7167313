#!/bin/bash
# Dev tool: the single-GPU evidence of a round in one gpurun call (tests, bench line, launch list, ncu captures).
# usage: tools/final_evidence.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-final}
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.txt 2>&1; tail -2 gpurun_out/${tag}_pytest.txt
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --quick --no-cpu --steps 6 --warmup 4 > gpurun_out/${tag}_launches.log 2>&1
ls -la gpurun_out/${tag}_*
