python tools/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1 && tail -4 gpurun_out/sanitize_plain.log &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_case.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
