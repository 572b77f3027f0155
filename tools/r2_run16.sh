for lib in cuda-raytracer_b200/libb2rt.so build/st24/libb2rt.so; do
  for tb in 20480 24576 28672 32768 40960; do
      printf "%-32s tb=%-6s host cfg2 : " $lib $tb; B2RT_BUILDER=host B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 32 --treelet-bytes $tb | tail -1
      printf "%-32s tb=%-6s host cfg3 : " $lib $tb; B2RT_BUILDER=host B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --treelet-bytes $tb | tail -1
  done
done
