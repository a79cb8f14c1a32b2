#!/usr/bin/env bash
# What to run first on a B200 after a change to the solver kernels (each block is one `gpurun -- '<block>'`; keep ncu out of
# multi-rank commands and off the cooperative-kernel path, see profiles/r01_launches_cylinder_v1.md).
#
#   bash examples/gpu_checklist.sh parity      # full parity suite (~12 min)
#   bash examples/gpu_checklist.sh long        # opt-in long tests (C++ channel example vs Orr-Sommerfeld, literal-rst Poiseuille)
#   bash examples/gpu_checklist.sh bench       # headline bench + phase table + per-kernel roofline at 24 k elements
#   bash examples/gpu_checklist.sh ab          # the environment-selectable kernel variants, A/B in one run
set -u
cd "$(dirname "$0")/.."
case "${1:-parity}" in
  parity)
    timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 ;;
  long)
    NLK_LONG_TESTS=1 timeout 900 python -m pytest tests/test_cpp_host.py tests/test_gpu_physics.py -m gpu -q 2>&1 | tail -5 ;;
  bench)
    NLK_PHASES=1 timeout 600 python bench.py --no-cpu 2>&1 | tail -14 | cut -c1-400
    timeout 300 python examples/kernel_bench.py --layers 12 --which 8,9,1,12,10,11,7,3 2>&1 | head -8
    timeout 300 python examples/kernel_bench.py --layers 50 --precond 4 --which 4,7,6 --nrep 10 2>&1 | head -3 ;;
  ab)
    for v in "" "NLK_AX8_LDSD=1" "NLK_OPDIV_MINB0=1 NLK_OPGRADT_MINB0=1 NLK_CONVECT_LB=0" "NLK_SWF8_MINB=6"; do
      echo "== $v"; env $v timeout 300 python examples/kernel_bench.py --layers 12 --which 8,0,11,10,3,7 --nrep 20 2>&1 | head -6
    done ;;
  *) echo "unknown block $1"; exit 2 ;;
esac
