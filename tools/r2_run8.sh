python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
sh tools/ab1.sh build/v3/libb2rt.so cuda-raytracer_b200/libb2rt.so
for cr in 512 1024 2048 4096; do export B2RT_CHUNK_RAYS=$cr; printf "chunk_rays %d " $cr; sh tools/ab1.sh cuda-raytracer_b200/libb2rt.so; done; unset B2RT_CHUNK_RAYS
for tb in 12288 16384 20480 22528; do printf "treelet %d cfg2: " $tb; python tools/profile_frame.py --frames 3 --spp 32 --treelet-bytes $tb | tail -1; done
for tb in 16384 20480 22528 24576 32768; do printf "treelet %d cfg3: " $tb; python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --treelet-bytes $tb | tail -1; done
