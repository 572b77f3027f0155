for lib in build/st24/libb2rt.so build/st32/libb2rt.so; do
  for tb in 28672 32768 40960 49152 57344; do
    for b in gpu host; do
      printf "%-24s tb=%-6s %-4s cfg2 : " $lib $tb $b; B2RT_BUILDER=$b B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 32 --treelet-bytes $tb | tail -1
      printf "%-24s tb=%-6s %-4s cfg3 : " $lib $tb $b; B2RT_BUILDER=$b B2RT_LIB=$lib python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --treelet-bytes $tb | tail -1
    done
  done
done
