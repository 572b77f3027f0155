for sl in "" "0.5,4,2" "1.0,4,2" "2.0,4,2" "1.0,3,3"; do
  printf "slice=%-10s cfg2/32spp : " "$sl"; B2RT_RENDER_SLICE=$sl timeout 120 python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "slice=%-10s cfg3/16spp : " "$sl"; B2RT_RENDER_SLICE=$sl timeout 120 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
