#!/bin/bash
# A/B of LDE generations on one B200: bench lines (no e2e, no CPU leg) per variant into gpurun_out/
set -u
mkdir -p gpurun_out
tag=${1:-ab}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${tag}_build.log 2>&1
for v in v4 nov4; do
  if [ $v = nov4 ]; then export TS_NO_V4=1; else unset TS_NO_V4; fi
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
done
unset TS_NO_V4
