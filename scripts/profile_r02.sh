#!/bin/bash
# Round-2 profiling pass (1 GPU, run under gpurun): ncu launch lists and full captures of the kernels that matter,
# compute-sanitizer racecheck / initcheck over the mbarrier ring, the hub merge, the 2-CTA GEMM and the new backward
# kernels.  Outputs land in gpurun_out/ (scratch); the summaries that are cited get copied into profiles/.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch lists (shares of the step; absolute times are cold-cache and serialised)
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench_c2.out 2>&1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02_launches_train.csv \
    python scripts/train_step_bench.py --scale 0.25 --iters 2 --profile 0 > gpurun_out/r02_launches_train.out 2>&1
# 2. full captures: edge kernel at C3 (d=400, RTE) and C5 (power-law, hub merge), backward GEMMs, edge backward
$NCU --set full --import-source on -k regex:k_edge_fwd -s 2 -c 1 -o gpurun_out/r02_edge_c3 \
    python scripts/quick_bench.py --config c3 --iters 2 --profile 0 > gpurun_out/r02_edge_c3.out 2>&1
$NCU --set full --import-source on -k regex:"k_edge_fwd|k_merge" -s 4 -c 2 -o gpurun_out/r02_edge_c5 \
    python scripts/quick_bench.py --config c5 --iters 2 --profile 0 > gpurun_out/r02_edge_c5.out 2>&1
$NCU --set full --import-source on -k regex:"k_lin_dx_tc|k_lin_dw_tc" -s 4 -c 4 -o gpurun_out/r02_bwd_gemm \
    python scripts/train_step_bench.py --scale 0.25 --iters 2 --profile 0 > gpurun_out/r02_bwd_gemm.out 2>&1
$NCU --set full --import-source on -k regex:"k_edge_bwd" -s 1 -c 1 -o gpurun_out/r02_edge_bwd \
    python scripts/train_step_bench.py --scale 0.25 --iters 2 --profile 0 > gpurun_out/r02_edge_bwd.out 2>&1
for f in r02_edge_c3 r02_edge_c5 r02_bwd_gemm r02_edge_bwd; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_ncu_raw.csv 2>/dev/null
done
# 3. sanitizers (small inputs): racecheck on shared-memory hazards (mbarrier ring, hub merge, epilogue staging),
#    initcheck on reads of uninitialised global memory
SEL="test_forward_matches_reference_golden or hub_split or tensor_core or backward_matches or sync_free or sharded_local"
compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py tests/test_gpu_bwd.py \
    -m gpu -q -x -k "$SEL" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
compute-sanitizer --tool initcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_bwd.py \
    -m gpu -q -x -k "$SEL" > gpurun_out/r02_sanitizer_initcheck.log 2>&1
tail -3 gpurun_out/r02_sanitizer_racecheck.log gpurun_out/r02_sanitizer_initcheck.log
