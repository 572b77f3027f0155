for ct in off 0.5 1.0 1.5 2.0 2.5 3.0; do
  for ml in 4 6 8; do
    if [ $ct = off ]; then unset B2RT_SAH_CT; else export B2RT_SAH_CT=$ct; fi
    printf "ct %-4s max_leaf %d cfg2: " $ct $ml; python tools/profile_frame.py --frames 3 --spp 32 --max-leaf $ml | tail -1
    printf "ct %-4s max_leaf %d cfg3: " $ct $ml; python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 --max-leaf $ml | tail -1
    printf "      "; python tools/count_work.py --max-leaf $ml
  done
done
