#!/bin/sh
# round 2: parity suite + A/B of library variants + one ncu capture of the traversal kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest.log
sh tools/ab.sh "$@" > gpurun_out/r2_ab.txt 2>&1
cat gpurun_out/r2_ab.txt
ncu --set full --clock-control none --import-source on -k regex:k_traverse -s 4 -c 4 -f -o gpurun_out/prof_r02 python tools/profile_frame.py --spp 32 > gpurun_out/ncu_r02.log 2>&1
tail -n 2 gpurun_out/ncu_r02.log
