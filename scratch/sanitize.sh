#!/bin/bash
# compute-sanitizer passes over the engine's kernels on small workloads (the reference has
# no sanitizer coverage).  usage: gpurun --timeout 900 -- 'bash scratch/sanitize.sh'
# Output: gpurun_out/sanitize_{memcheck,racecheck,initcheck}.log (+ a one-line summary each in sanitize.log).
mkdir -p gpurun_out
SEL='test_gpu_zz_fixtures.py::test_engine_replays_reference_kernel_fixture'
: > gpurun_out/sanitize.log
for tool in memcheck racecheck initcheck; do
    timeout 280 compute-sanitizer --tool $tool --error-exitcode 86 --log-file gpurun_out/sanitize_$tool.log \
        python -m pytest "tests/$SEL" -x -q -m gpu -k "tiny or scint or wires" > gpurun_out/sanitize_${tool}_pytest.log 2>&1
    echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)" >> gpurun_out/sanitize.log
done
cat gpurun_out/sanitize.log
