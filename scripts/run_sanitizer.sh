# GPU box: compute-sanitizer over the small end-to-end run
{
echo '$ compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py'
timeout 270 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py 2>&1 | grep -E "sanitize run|ERROR SUMMARY|Invalid|Error" | head -20
echo '$ compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py'
timeout 170 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/sanitize_small.py 2>&1 | grep -E "sanitize run|RACECHECK SUMMARY|hazard|Error" | head -20
} > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
