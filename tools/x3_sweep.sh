#!/bin/bash
# A/B sweeps of the 3xTF32 kernels on one GPU (debug helper; prints it/s and the per-phase ms)
run() { timeout 200 python bench.py --steps 4 --warmup 2 --t-scale 0.25 "$@" --no-e2e --no-cpu-baseline --no-fp32-grade 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.2f it/s' % d['value'], d['roofline']['kernel_ms_per_step'])"; }
for sub in 8 16 0; do echo "== main X3_SUB=$sub"; CMF_X3_SUB=$sub run --precision tf32x3; done
for v in nosetmax r56; do for sub in 8 0; do echo "== $v X3_SUB=$sub"; CMF_B200_LIB=$PWD/build/variants/lib_$v.so CMF_X3_SUB=$sub run --precision tf32x3; done; done
echo "== tf32 main"; run
for sub in 8 16; do echo "== traj X3_SUB=$sub"; CMF_X3_SUB=$sub timeout 300 python tools/trajectory_errors.py tf32x3,tf32x3g A,mid,B,tc_l70; done
