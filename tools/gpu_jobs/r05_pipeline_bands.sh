#!/bin/bash
# round 2 (session 5): pipelined copies: parity tests, then the band count A/B on the bench workload
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout 600 -x) > gpurun_out/r05_pytest_pipeline.log 2>&1; echo "pipeline rc=$?"; tail -5 gpurun_out/r05_pytest_pipeline.log
for nb in 32 64; do
HMI_PIPELINE_BANDS=$nb timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bands $nb value %.1f e2e %.1f ratio %.3f parity %s' % (d['value'], d['e2e']['value'], d['e2e']['value'] / d['value'], d['parity']['pass']))"
done | tee gpurun_out/r05_tune_pipeline_bands_merged_ramps.log
