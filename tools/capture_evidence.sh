#!/bin/bash
# ncu evidence for profiles/: full captures of the dominant kernels + a bounded launch list of the bench command.
# Run on the GPU box: bash tools/capture_evidence.sh <tag>
TAG=${1:-r1b}
O=gpurun_out
mkdir -p $O
NCU="ncu --profile-from-start off --clock-control none"
U2_PROBE_STEPS=1 timeout 240 $NCU --set full --import-source on -k regex:dlinear -s 5 -c 2 -f -o $O/${TAG}_dlinear_chain_full \
  python tools/decode_probe.py > $O/${TAG}_ncu_dlinear.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:fa_fwd -c 1 -f -o $O/${TAG}_fa_fwd2_full \
  python tools/vision_probe.py > $O/${TAG}_ncu_fa.log 2>&1
timeout 200 $NCU --set full --import-source on -k regex:gemm_bf16 -s 1 -c 4 -f -o $O/${TAG}_gemm_full \
  python tools/vision_probe.py > $O/${TAG}_ncu_gemm.log 2>&1
U2_PROFILE_TIMED=1 timeout 500 $NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/${TAG}_bench_cfg3_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_under_ncu.log 2>&1
ls -la $O | tail -12
