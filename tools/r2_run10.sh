python -m pytest tests -m gpu -q 2>&1 | tail -n 8
# launch list (time + DRAM bytes per launch) of one 32-spp frame of the cfg3 stand-in and of cfg2
python tools/profile_frame.py --spp 32 --subdivide 1 --width 1920 --height 1080 --overlap > gpurun_out/plain_cfg3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02_cfg3.csv python tools/profile_frame.py --spp 32 --subdivide 1 --width 1920 --height 1080 > gpurun_out/ncu_l_cfg3.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02_cfg3.csv
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02_cfg2.csv python tools/profile_frame.py --spp 64 > gpurun_out/ncu_l_cfg2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_r02_cfg2.csv
