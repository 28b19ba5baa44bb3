CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 600 $CS --tool $tool --print-limit 20 python tools/sanitize_case.py 6 2 7 5 > gpurun_out/sanitizer_${tool}_cfg1.log 2>&1
  tail -4 gpurun_out/sanitizer_${tool}_cfg1.log
done
timeout 900 $CS --tool memcheck --print-limit 20 python tools/sanitize_case.py 250 4 40 60 > gpurun_out/sanitizer_memcheck_cfg2slice.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck_cfg2slice.log
timeout 900 $CS --tool racecheck --print-limit 20 python tools/sanitize_case.py 120 2 40 40 > gpurun_out/sanitizer_racecheck_cfg2slice.log 2>&1
tail -4 gpurun_out/sanitizer_racecheck_cfg2slice.log
