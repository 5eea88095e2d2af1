#!/bin/bash
# ncu captures of the main kernels, summarised ON the GPU box (the raw reports exceed what gpurun copies back):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/profile_all.sh'
# leaves profiles-style summaries in gpurun_out/r2_*.md and the launch list of the bench command in gpurun_out/launches_r2.csv
set -u
mkdir -p gpurun_out /tmp/prof
N="ncu --set full --clock-control none --import-source on -f"
$N -k regex:meshSample -s 1 -c 1 -o /tmp/prof/mesh python tools/ncu_target.py mesh > gpurun_out/ncu_mesh.log 2>&1
$N -k regex:fitKernel -s 2 -c 2 -o /tmp/prof/fitjit python tools/ncu_target.py fitjit > gpurun_out/ncu_fitjit.log 2>&1
$N -k regex:queryKernel -s 2 -c 1 -o /tmp/prof/query python tools/ncu_target.py query > gpurun_out/ncu_query.log 2>&1
$N -k regex:"schedRound|schedSelect|schedIngest" -s 12 -c 6 -o /tmp/prof/sched python tools/ncu_target.py sched > gpurun_out/ncu_sched.log 2>&1
timeout 240 $N -k regex:"cgKernel|faceEnum|faceEmit" -c 4 -o /tmp/prof/cont python tools/ncu_target.py continuity > gpurun_out/ncu_cont.log 2>&1
$N -k regex:"meshPseudo|meshLevelKeys|meshObb|meshRefit" -c 4 -o /tmp/prof/meshbuild python tools/mesh_create_time.py > gpurun_out/ncu_meshbuild.log 2>&1
python tools/summarize_ncu.py report /tmp/prof/mesh.ncu-rep gpurun_out/r2_mesh_sample_kernel.md "Round 2: meshSampleKernel (with the tail phase), first round of the 870 000-triangle config (4096 coarse fits, 3.0 M samples)"
python tools/summarize_ncu.py report /tmp/prof/fitjit.ncu-rep gpurun_out/r2_fit_kernel_jit.md "Round 2: fitKernel<D,false> specialised at run time (NVRTC, parameters in constant memory) for the C2 program, synthetic frontier p=2"
python tools/summarize_ncu.py report /tmp/prof/query.ncu-rep gpurun_out/r2_query_kernel.md "Round 2: queryKernel, 16.7 M uniform points on the C2 tree"
python tools/summarize_ncu.py report /tmp/prof/sched.ncu-rep gpurun_out/r2_scheduler_kernels.md "Round 2: device-resident scheduler kernels of a C2 build (rounds 3-5)"
python tools/summarize_ncu.py report /tmp/prof/cont.ncu-rep gpurun_out/r2_continuity_kernels.md "Round 2: continuity kernels on csg_cont (face enumeration, face emission, CG)"
python tools/summarize_ncu.py report /tmp/prof/meshbuild.ncu-rep gpurun_out/r2_mesh_build_kernels.md "Round 2: kernels of hpsdf_mesh_create, 870 000 triangles"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --only query --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_r2.csv gpurun_out/r2_launches.md "Round 2: kernel launches of python bench.py --steps 2 --warmup 3 --only query --no-cpu-baseline"
ls -la gpurun_out/*.md
