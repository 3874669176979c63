#!/bin/bash
# ncu --set full of the multi-pick FPS launches of one lone forward + a fresh launch list (summarised on the box).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k "regex:fps_multi" -s 3 -c 3 -o gpurun_out/r02_prof_fps_multi python scripts/one_forward.py tc serial > gpurun_out/r02_prof_fps_multi.log 2>&1
python scripts/ncu_summary.py full gpurun_out/r02_prof_fps_multi.ncu-rep > gpurun_out/r02_prof_fps_multi_full.txt
rm -f gpurun_out/r02_prof_fps_multi.ncu-rep
cat gpurun_out/r02_prof_fps_multi_full.txt | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_forward.csv python scripts/one_forward.py tc serial > /dev/null 2>&1
