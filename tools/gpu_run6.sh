#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:"k_cubic_rgba|k_box_rgba" -s 8 -c 3 -o gpurun_out/r1h_prof_mip_levels_survey -f python bench.py --workload c3 --steps 4 --c3-data survey > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_box_rgba" -s 4 -c 1 -o gpurun_out/r1h_prof_box_random -f python bench.py --workload c3 --steps 4 --c3-data random > /dev/null 2>&1
ls -la gpurun_out | grep r1h
