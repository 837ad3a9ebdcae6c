#!/bin/bash
# One GPU call that refreshes the evidence under profiles/ for a round (outputs land in gpurun_out/<tag>_*):
#   launch lists (cold and warm caches) of the bench workload's frame, full-set captures of its kernels on one C4 view,
#   on the C2 frame and on the Sponza frame (binner / tile raster under load), then the contract bench line.
# usage: bash tools/profile_round.sh <tag>
tag=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c4.csv python tools/profile_frame.py c4_views 6 0 binned > gpurun_out/${tag}_prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c4_warm.csv python tools/profile_frame.py c4_views 6 0 binned >> gpurun_out/${tag}_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_mesh_setup|k_bin_scatter|k_tile_raster|k_resolve|k_fb_detile" -s 4 -c 5 -o gpurun_out/${tag}_full_c4 -f python tools/profile_frame.py c4_views 3 0 binned >> gpurun_out/${tag}_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_mesh_setup|k_frame_begin" -s 2 -c 2 -o gpurun_out/${tag}_full_c2 -f python tools/profile_frame.py c2_grid 3 0 binned >> gpurun_out/${tag}_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_mesh_setup|k_bin_scatter|k_tile_raster|k_resolve" -s 4 -c 4 -o gpurun_out/${tag}_full_sponza -f python tools/profile_frame.py c1_sponza 3 0 binned >> gpurun_out/${tag}_prof.log 2>&1
tail -3 gpurun_out/${tag}_prof.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err; tail -c 400 gpurun_out/${tag}_bench_n1.json
