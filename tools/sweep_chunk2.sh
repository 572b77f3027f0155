#!/bin/sh
# adaptive chunk size of the deeper levels: chunks per resident CTA x smallest chunk
for cpc in 0 2 4 8; do for cmin in 128 256 512; do
  [ "$cpc" = 0 ] && [ "$cmin" != 256 ] && continue
  echo "chunks_per_cta $cpc chunk_min $cmin"
  B2RT_CHUNKS_PER_CTA=$cpc B2RT_CHUNK_MIN=$cmin sh tools/ab1.sh cuda-raytracer_b200/libb2rt.so
done; done
