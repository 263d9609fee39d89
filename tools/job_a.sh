#!/bin/bash
# scratch GPU job: parity tests, chunk-policy / skew A/B, e2e breakdown, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
for v in skew noskew; do
  if [ $v = noskew ]; then export HMI_B200_LIB=$PWD/heightmap_interpolation_b200/csrc/variants/libhmi_b200_noskew.so; else unset HMI_B200_LIB; fi
  echo "== variant $v n=4096" >> gpurun_out/tune_ab.log
  timeout 600 python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 2,3,4,6,8 --chunks=-1,0 --vecs 2 >> gpurun_out/tune_ab.log 2>&1
  echo "== variant $v n=16384" >> gpurun_out/tune_ab.log
  timeout 600 python tools/tune.py --n 16384 --methods ccst,harmonic --dtypes f64 --kernels stream --tbs 3,4,6 --chunks=-1,0,256 --vecs 2 >> gpurun_out/tune_ab.log 2>&1
done
unset HMI_B200_LIB
timeout 300 python tools/tune.py --n 4096 --methods tv,amle --dtypes f64,f32 --kernels stream --chunks=-1,0 >> gpurun_out/tune_ab.log 2>&1
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log
