for lib in cuda-raytracer_b200/libb2rt.so build/st20/libb2rt.so build/st24/libb2rt.so; do
  for tb in 0 32768; do
    printf "%-32s tb=%-6s cfg2 : " $lib $tb; B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 32 --treelet-bytes $tb | tail -1
    printf "%-32s tb=%-6s cfg3 : " $lib $tb; B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --treelet-bytes $tb | tail -1
  done
done
printf "host builder cfg2: "; B2RT_BUILDER=host python tools/profile_frame.py --frames 3 --spp 32 | tail -1
printf "host builder cfg3: "; B2RT_BUILDER=host python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 | tail -1
