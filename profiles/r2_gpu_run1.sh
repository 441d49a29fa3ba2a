#!/usr/bin/env bash
# Round 2, GPU call 1: new kernel quick check, full GPU suite, bench (parity gate), sanitizer probe.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/r2a_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q -k "reduced_system_kernels_and_modes_agree" > $O/r2a_rows_quick.log 2>&1
echo "rows quick rc=$?" | tee -a $O/r2a_rows_quick.log
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -k "not cfg4_full_size and not cfg5_full_size" > $O/r2a_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2a_tests.log
tail -5 $O/r2a_tests.log
timeout 900 python bench.py > $O/r2a_bench.json 2> $O/r2a_bench.err
echo "bench rc=$?" | tee -a $O/r2a_bench.err
timeout 300 python bench.py --no-configs --no-c64 --steps 10 > $O/r2a_bench_spb5.json 2>> $O/r2a_bench.err
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python profiles/sanitize_case.py rows_tma > $O/r2a_san_${tool}_rows_tma.txt 2>&1
  echo "exit $?" >> $O/r2a_san_${tool}_rows_tma.txt
done
compute-sanitizer --version > $O/r2a_san_version.txt 2>&1
tail -c 1500 $O/r2a_bench.json
