#!/bin/sh
# adaptive chunk size of the deeper levels, larger upper bound
for cr in 1024 2048 4096; do for cpc in 2 3 4; do
  echo "chunk_rays $cr chunks_per_cta $cpc"
  B2RT_CHUNK_RAYS=$cr B2RT_CHUNKS_PER_CTA=$cpc sh tools/ab1.sh cuda-raytracer_b200/libb2rt.so
done; done
