# Round-end evidence on one B200 (run through gpurun; everything lands in gpurun_out/, the big .ncu-rep files are
# summarised on the box and removed: gpurun copies back at most 64 MiB).
set -x
if [ -z "$SKIP_PYTEST" ]; then python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log; fi
python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err; tail -c 300 gpurun_out/f_bench_default.err
python tools/trajectory_errors.py fp32,tf32x3,tf32x3g,tf32,tf32g > gpurun_out/f_traj.log 2>&1; tail -3 gpurun_out/f_traj.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-other-configs > gpurun_out/f_ncu_list.log 2>&1
python tools/ncu_summary.py list gpurun_out/f_launches.csv > gpurun_out/f_launch_shares.txt 2>&1
for mode in tf32x3 tf32; do
  ncu --set full --clock-control none -k regex:tc_ -s 12 -c 10 -o gpurun_out/f_$mode python bench.py --precision $mode --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-other-configs --no-tf32 > gpurun_out/f_ncu_$mode.log 2>&1
  python tools/ncu_summary.py full gpurun_out/f_$mode.ncu-rep > gpurun_out/f_ncu_kernels_$mode.txt 2>&1
  rm -f gpurun_out/f_$mode.ncu-rep
done
ncu --set full --clock-control none -k regex:"tc_hterms|tc_recon_x3|tc_wterms_x3" -s 8 -c 4 -o gpurun_out/f_B python bench.py --config B --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-peak --no-tf32 --no-other-configs > gpurun_out/f_ncu_B.log 2>&1
python tools/ncu_summary.py full gpurun_out/f_B.ncu-rep > gpurun_out/f_ncu_kernels_B.txt 2>&1
rm -f gpurun_out/f_B.ncu-rep
ls -la gpurun_out/f_*
