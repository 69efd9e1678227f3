set -x
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_probe.py > gpurun_out/r02d_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r02d_sanitizer_memcheck.log
timeout 200 compute-sanitizer --tool racecheck python tools/sanitize_probe.py > gpurun_out/r02d_sanitizer_racecheck.log 2>&1; tail -4 gpurun_out/r02d_sanitizer_racecheck.log
