for lib in build/r1/libb2rt.so cuda-raytracer_b200/libb2rt.so build/dm8/libb2rt.so build/dm16/libb2rt.so build/dm24/libb2rt.so; do printf "%-34s " $lib; B2RT_LIB=$lib python tools/count_work.py; done
sh tools/ab1.sh cuda-raytracer_b200/libb2rt.so build/dm8/libb2rt.so build/dm16/libb2rt.so build/dm24/libb2rt.so
for ml in 2 3 4 6 8; do printf "max_leaf %d : " $ml; python tools/profile_frame.py --frames 3 --spp 32 --max-leaf $ml | tail -1; python tools/count_work.py --max-leaf $ml; done
