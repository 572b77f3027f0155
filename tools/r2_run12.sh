# device builder with PLOC vs LBVH vs the host SAH builder: parity suite on the device-built tree, frame times, build times
B2RT_BUILDER=gpu python -m pytest tests -m gpu -x -q -k "radiance or device_built or waves or brute" 2>&1 | tail -n 4
for cfg in "host ploc" "gpu ploc" "gpu lbvh"; do
  set -- $cfg
  printf "builder=%-5s binary=%-5s cfg2/32spp : " $1 $2; B2RT_BUILDER=$1 B2RT_GPU_BINARY=$2 python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
  printf "builder=%-5s binary=%-5s cfg3/16spp : " $1 $2; B2RT_BUILDER=$1 B2RT_GPU_BINARY=$2 python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
done
B2RT_LARGE_PRIM=0 B2RT_BUILDER=gpu python tools/profile_frame.py --frames 3 --spp 32 2>&1 | tail -1
B2RT_LARGE_PRIM=0 B2RT_BUILDER=gpu python tools/profile_frame.py --frames 3 --spp 16 --subdivide 1 --width 1920 --height 1080 2>&1 | tail -1
python tools/bench_build.py --tris 10000000 --rays 4194304 2>&1 | tail -n 8
B2RT_GPU_BINARY=lbvh python tools/bench_build.py --tris 10000000 --rays 4194304 2>&1 | grep gpu | tail -n 2
B2RT_VERBOSE=1 B2RT_BUILDER=gpu python tools/profile_frame.py --frames 1 --spp 4 --subdivide 1 --width 1920 --height 1080 2>&1 | grep "gpu build" | tail -12
