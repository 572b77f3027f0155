#!/bin/sh
# CTAs per SM of the scheduler's histogram / scatter kernels
for c in 4 8; do for sc in 2 4 6 8; do
  echo "count_ctas $c scatter_ctas $sc"
  B2RT_COUNT_CTAS=$c B2RT_SCATTER_CTAS=$sc sh tools/ab1.sh cuda-raytracer_b200/libb2rt.so
done; done
