#!/bin/bash
# One GPU call that refreshes everything under profiles/ for a round: launch list, full-set captures (C2 frame, C4 draw).
# usage: bash tools/profile_round.sh <tag>      (outputs land in gpurun_out/<tag>_*)
tag=${1:-r01b}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2_binned.csv python tools/profile_frame.py 8 > gpurun_out/${tag}_prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2_binned_warm.csv python tools/profile_frame.py 8 >> gpurun_out/${tag}_prof.log 2>&1
ncu --set full --clock-control none --import-source on -s 11 -c 6 -o gpurun_out/${tag}_full_c2 -f python tools/profile_frame.py 4 binned c2 >> gpurun_out/${tag}_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_mesh_setup|k_bin_scatter|k_tile_raster" -s 3 -c 3 -o gpurun_out/${tag}_full_c4 -f python tools/profile_frame.py 3 binned c4 >> gpurun_out/${tag}_prof.log 2>&1
tail -3 gpurun_out/${tag}_prof.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err; tail -c 600 gpurun_out/${tag}_bench_n1.json
