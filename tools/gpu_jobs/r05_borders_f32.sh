#!/bin/bash
# round 2 (session 5): throughput of the zero-padded border rule (SciPy convolvers) on the streaming kernels vs the
# generic kernel; fp32 CCST 3 vs 4 fused sweeps after the steady-state loop change
mkdir -p gpurun_out
{
python tools/tune.py --n 4096 --methods harmonic --dtypes f64,f32 --kernels stream,generic --tbs 1,6 --chunks 0 --vecs 2 --sweeps 96 --border zero --orientation -1
python tools/tune.py --n 4096 --methods ccst --dtypes f64,f32 --kernels stream,generic --tbs 1,3 --chunks 0 --vecs 2 --sweeps 96 --border zero --orientation -1
python tools/tune.py --n 4096 --methods tv,amle --dtypes f64,f32 --kernels stream,generic --tbs 1,2 --chunks 0 --vecs 2 --sweeps 96 --border zero --orientation -1
python tools/tune.py --n 4096 --methods harmonic,ccst --dtypes f64 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 96
} > gpurun_out/r05_tune_zero_border_stream_vs_generic.log 2>&1
cat gpurun_out/r05_tune_zero_border_stream_vs_generic.log
{
for n in 4096 8192 16384 32768; do
python tools/tune.py --n $n --methods ccst --dtypes f32 --kernels stream --tbs 3,4 --chunks 0 --vecs 2 --sweeps 48 --mask random
python tools/tune.py --n $n --methods harmonic --dtypes f32 --kernels stream --tbs 6,8 --chunks 0 --vecs 2 --sweeps 48 --mask random
done
} > gpurun_out/r05_tune_f32_temporal_block.log 2>&1
cat gpurun_out/r05_tune_f32_temporal_block.log
