#!/usr/bin/env bash
# Round 2, GPU call 10 (re-run as r2o after the time-parallel mode and NVTX): full GPU suite + bench on the chain_duo build
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 > $O/r2o_tests.log 2>&1
echo "tests rc=$?" | tee -a $O/r2o_tests.log
tail -6 $O/r2o_tests.log
timeout 900 python bench.py > $O/r2o_bench.json 2> $O/r2o_bench.err
echo "bench rc=$?" | tee -a $O/r2o_bench.err
tail -c 600 $O/r2o_bench.err
python -c "
import json;d=json.load(open('$O/r2o_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['parity']['ok'], d['clocks'])"
