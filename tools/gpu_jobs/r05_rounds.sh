#!/bin/bash
# round 2 (session 5): does a launch of few waves gain from more, shorter rounds (load balance) than the picker's cost model chooses?
mkdir -p gpurun_out
for r in 1 2 3 4 6 8; do
export HMI_PICK_MIN_ROUNDS=$r
echo "== min rounds $r"
HMI_DEBUG=1 python tools/tune.py --rows 4096 --n 32768 --methods harmonic --dtypes f32 --kernels stream --tbs 8 --chunks 0 --vecs 2 --sweeps 96 --mask random 2>&1 | grep -v "norm=1" | sort -u
HMI_DEBUG=1 python tools/tune.py --rows 4096 --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 96 --mask random 2>&1 | grep -v "norm=1" | sort -u
HMI_DEBUG=1 python tools/tune.py --rows 4096 --n 32768 --methods ccst --dtypes f32 --kernels stream --tbs 4 --chunks 0 --vecs 2 --sweeps 96 --mask random 2>&1 | grep -v "norm=1" | sort -u
HMI_DEBUG=1 python tools/tune.py --rows 512 --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 96 --mask random 2>&1 | grep -v "norm=1" | sort -u
HMI_DEBUG=1 python tools/tune.py --n 4096 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2 --sweeps 96 2>&1 | grep -v "norm=1" | sort -u
done > gpurun_out/r05_tune_picker_min_rounds.log 2>&1
cat gpurun_out/r05_tune_picker_min_rounds.log
