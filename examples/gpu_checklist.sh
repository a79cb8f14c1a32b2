#!/usr/bin/env bash
# What to run first on a B200 after a change to the solver kernels (each block is one `gpurun -- '<block>'`; keep ncu out of
# multi-rank commands and off the cooperative-kernel path, see profiles/r01_launches_cylinder_v1.md).
#
#   bash examples/gpu_checklist.sh parity      # full parity suite (~6 min)
#   bash examples/gpu_checklist.sh cg3         # validate + A/B the opt-in three-field Helmholtz PCG (never run so far)
#   bash examples/gpu_checklist.sh long        # tests written after the round-1 GPU budget was spent (C++ example, literal-rst Poiseuille)
#   bash examples/gpu_checklist.sh bench       # default bench + phase table + per-kernel roofline at 32 k elements
set -u
cd "$(dirname "$0")/.."
case "${1:-parity}" in
  parity)
    timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 ;;
  cg3)
    # streamed path on the small parity cases (NLK_NO_CGP), then the size where the streamed path is the default
    NLK_CG3=1 NLK_NO_CGP=1 timeout 600 python -m pytest tests/test_gpu_exptA.py tests/test_gpu_kernels.py tests/test_gpu_properties.py -m gpu -q -x 2>&1 | tail -5
    for v in 0 1; do
      if [ $v = 1 ]; then export NLK_CG3=1; else unset NLK_CG3; fi
      NLK_PHASES=1 timeout 300 python bench.py --no-cylinder --cpu-steps 0 2>&1 | grep -E "Helmholtz|ms_per_step" | cut -c1-200
    done ;;
  long)
    NLK_LONG_TESTS=1 timeout 900 python -m pytest tests/test_cpp_host.py tests/test_gpu_physics.py -m gpu -q 2>&1 | tail -5 ;;
  bench)
    NLK_PHASES=1 timeout 400 python bench.py 2>&1 | tail -14 | cut -c1-400
    timeout 300 python examples/kernel_bench.py --layers 16 2>&1 | head -6
    timeout 200 python examples/kernel_bench.py --layers 6 --precond 4 --which 6,7,4 --nrep 50 2>&1 | head -3 ;;
  *) echo "unknown block $1"; exit 2 ;;
esac
