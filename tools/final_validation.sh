#!/bin/bash
# Round-end validation on the GPU box: full GPU test suite, smoke, the bench lines and the preprocessing probe.
TAG=${1:-r1_final}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $O/${TAG}_bench_cfg3.json 2> $O/${TAG}_bench_cfg3.err; echo "bench cfg3 rc=$?"; cut -c1-600 $O/${TAG}_bench_cfg3.json
timeout 300 python bench.py --workload cfg2 --no-cpu-baseline > $O/${TAG}_bench_cfg2.json 2> $O/${TAG}_bench_cfg2.err; echo "bench cfg2 rc=$?"; cut -c1-300 $O/${TAG}_bench_cfg2.json
timeout 120 python tools/preprocess_probe.py 2>&1 | tail -2
