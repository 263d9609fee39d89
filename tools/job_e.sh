#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu --timeout 180 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for n in 4096 16384; do
timeout 600 python tools/tune.py --n $n --methods tv,amle --dtypes f64,f32 --kernels stream --chunks=-1,0 --vecs 2 >> gpurun_out/tune_nl2.log 2>&1
done
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log | cut -c1-1500
