#!/bin/sh
for wave in 4194304 6300000 7900000 9500000 12600000 16000000 25200000; do printf "wave=%s : " $wave; timeout 80 python tools/profile_frame.py --spp 64 --frames 2 --wave $wave | tail -1; done
