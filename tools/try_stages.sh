set -e
for st in 6 10; do
  U2_NVCC_DEFINES="U2_DL_STAGES128=$st" python u2tokenizer_b200/build.py --force > /dev/null
  echo "== stages $st"
  U2_L2_LOOKAHEAD=0 U2_L2_NEXT=-1 timeout 100 python tools/chain_probe.py 2>&1 | tail -5 | head -1
  U2_L2_LOOKAHEAD=0 U2_L2_NEXT=-1 timeout 300 python tools/profile_phases.py --workload cfg3 2>&1 | tail -2
  U2_L2_LOOKAHEAD=0 U2_L2_NEXT=20 timeout 300 python tools/profile_phases.py --workload cfg3 2>&1 | tail -2
done
