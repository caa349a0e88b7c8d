#!/bin/bash
# Round-2 profile collection, run ON the GPU box (under gpurun): launch lists + `ncu --set full` captures of the hot
# kernels.  Only text digests are left under gpurun_out/ (the .ncu-rep files exceed the 64 MiB copy-back limit):
#   r02_launches_bench.csv / r02_launches_mining.csv   per-launch device time (cold cache, serialised: compare SHARES)
#   r02_ncu_summary.md                                  tools/ncu_summary.py over every capture
#   r02_sass_hot_<kernel>.txt                           tools/ncu_sass_hot.py: opcode mix, stall reasons, hottest lines
set -u
T=/tmp/en_prof; mkdir -p $T gpurun_out build
# developer binaries (build/ is git-ignored): the PCIe / kernel overlap probe and the library variant with trace stamps
[ -x build/pipe_probe ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pipe_probe.cu -o build/pipe_probe
[ -f build/lib_trace.so ] || python -c "from embeddingnet_b200 import build as b; b.build_variant('build/lib_trace.so', ['-DEN_FIN_TRACE'])"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --skip-knn --skip-cpu > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_mining.csv \
    python tools/profile_step.py mining > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_smallq.csv \
    python tools/profile_step.py smallq > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"dist_gemm|batch_hard_finalize|split_planes" -s 8 -c 4 -o $T/bh \
    python tools/profile_step.py triplet > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"pair_tc_kernel|pair_finish|collect_positives" -c 8 -o $T/pair \
    python tools/profile_step.py pairbwd > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"pair_tc_kernel" -c 1 -o $T/ba64 \
    python tools/profile_step.py ba64 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_ba64.csv \
    python tools/profile_step.py ba64 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"knn_smallq|knn_stream" -c 6 -o $T/smallq \
    python tools/profile_step.py smallq > /dev/null 2>&1
$NCU --set full -k regex:"l2norm|row_dist" -s 4 -c 4 -o $T/rowwise python tools/profile_step.py rowwise > /dev/null 2>&1
# (-k matches the base name, which carries no template arguments: match the demangled name for the epilogue type)
$NCU --set full --import-source on --kernel-name-base demangled -k regex:"EpMine" -c 2 -o $T/mine python tools/profile_step.py mining > /dev/null 2>&1
python tools/ncu_summary.py $T/bh.ncu-rep $T/pair.ncu-rep $T/ba64.ncu-rep $T/smallq.ncu-rep $T/rowwise.ncu-rep $T/mine.ncu-rep \
    > gpurun_out/r02_ncu_summary.md 2> gpurun_out/r02_ncu_summary.err
hot() {  # <rep> <kernel id (1-based)> <name>
  ncu -i $T/$1.ncu-rep --page source --csv --kernel-id :::$2 > $T/src.csv 2>/dev/null
  python tools/ncu_sass_hot.py $T/src.csv 25 > gpurun_out/r02_sass_hot_$3.txt 2>&1
}
hot bh 2 bh_gemm; hot bh 3 bh_finalize_fast; hot bh 4 bh_finalize_slow
hot pair 2 pair_tc_batch_all; hot pair 6 pair_tc_contrastive
hot smallq 3 knn_smallq_q8; hot mine 1 mine_count; hot ba64 1 pair_tc_batch_all_big
EMBEDDINGNET_B200_LIB=build/lib_trace.so python tools/trace_bh.py > gpurun_out/r02_trace_bh_step.txt 2>&1
python tools/time_bh.py > gpurun_out/r02_time_bh.txt 2>&1
python tools/time_ba64.py > gpurun_out/r02_time_batch_all.txt 2>&1
./build/pipe_probe 3 > gpurun_out/r02_pipe_probe.txt 2>&1
ls -la $T gpurun_out | tail -30
