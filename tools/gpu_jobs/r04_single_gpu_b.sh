#!/bin/bash
# round 2 (session 4), single GPU: cells-per-lane A/B (fp64 V = 2 / 4), staging A/B with ncu summaries made on the box
# (the reports stay in /tmp: five of them were over the 64 MiB limit of gpurun_out last time)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "converged" > gpurun_out/r04_pytest_converged.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r04_pytest_converged.log
{
python tools/tune.py --n 4096 --methods ccst --dtypes f64 --kernels stream --tbs 2,3 --chunks 0 --vecs 2,4 --sweeps 96
python tools/tune.py --n 4096 --methods harmonic --dtypes f64 --kernels stream --tbs 4,6,7 --chunks 0 --vecs 2,4 --sweeps 96
python tools/tune.py --n 16384 --methods ccst --dtypes f64 --kernels stream --tbs 2,3 --chunks 0 --vecs 2,4 --sweeps 48 --mask random
python tools/tune.py --n 16384 --methods harmonic --dtypes f64 --kernels stream --tbs 6,7 --chunks 0 --vecs 2,4 --sweeps 48 --mask random
python tools/tune.py --n 32768 --methods ccst --dtypes f64 --kernels stream --tbs 3 --chunks 0 --vecs 2,4 --sweeps 24 --mask random
python tools/tune.py --n 32768 --methods ccst,harmonic --dtypes f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 24 --mask random
} > gpurun_out/r04_tune_vec.log 2>&1
echo "tune vec rc=$?"; grep GLUPS gpurun_out/r04_tune_vec.log
{
for tma in 0 2; do
python tools/tune.py --n 4096 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 96 --tma $tma
python tools/tune.py --n 16384 --methods ccst,harmonic --dtypes f64,f32 --kernels stream --tbs 3,6 --chunks 0 --vecs 2 --sweeps 48 --mask random --tma $tma
done
} > gpurun_out/r04_tune_staging_tma2d.log 2>&1
echo "tune tma rc=$?"; grep GLUPS gpurun_out/r04_tune_staging_tma2d.log
cap() {  # name, prof_one args...
  name=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 2 -c 1 -f -o /tmp/$name python tools/prof_one.py "$@" > /tmp/$name.log 2>&1; echo "ncu $name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  python tools/ncu_opmix.py /tmp/$name.ncu-rep 1 2>&1 | head -45 >> gpurun_out/${name}_summary.txt
}
cap r04_ncu_ccst_f64_k3_32768 --n 32768 --method ccst --dtype f64 --tb 3 --sweeps 12 --mask random
cap r04_ncu_ccst_f64_k3_32768_vec4 --n 32768 --method ccst --dtype f64 --tb 3 --sweeps 12 --mask random --vec 4
cap r04_ncu_ccst_f64_k3_16384_tma0 --n 16384 --method ccst --dtype f64 --tb 3 --sweeps 12 --mask random
cap r04_ncu_ccst_f64_k3_16384_tma2 --n 16384 --method ccst --dtype f64 --tb 3 --sweeps 12 --mask random --tma 2
ncu -i /tmp/r04_ncu_ccst_f64_k3_32768.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r04_ncu_ccst_f64_k3_32768_source.csv.gz
ncu -i /tmp/r04_ncu_ccst_f64_k3_32768_vec4.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r04_ncu_ccst_f64_k3_32768_vec4_source.csv.gz
ls -la gpurun_out
