// ORACLE / TEST INFRASTRUCTURE -- second reported baseline: the reference's OWN CUDA renderer
// (src/cudaRenderer.cu, compiled unmodified for sm_100a from where it lies under /root/reference) driven
// headless.  This file is ours: it replaces src/cudaMain.cpp + src/display.cpp (GLUT window) with the same
// call sequence (cudaMain.cpp:88-99: allocOutputImage -> loadScene -> setup, then display.cpp:120-136:
// render() per displayed frame) and times the frames with a host clock around cudaDeviceSynchronize.
//
//   ref_cuda_render <scene.dae> [frames=16] [warmup=3] [size=512]
//
// The reference renders SAMPLES_PER_PIXEL (2) samples per render() on a square power-of-two image with a
// fixed script of 8 traversal passes (primary, 2 + 2 + 1 shadow passes, 2 scene bounces:
// cudaRenderer.cu:2499-2534), so a frame traces at most 8 * size^2 * 2 rays.  Prints one JSON line.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "cudaRenderer.h"

// the reference's error dialog (src/error_dialog.cpp) needs GLUT; collada.cpp only calls this entry point
void showError(std::string msg, bool fatal) {
  fprintf(stderr, "reference error: %s\n", msg.c_str());
  if (fatal) exit(1);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s scene.dae [frames] [warmup] [size]\n", argv[0]); return 2; }
  const int frames = argc > 2 ? atoi(argv[2]) : 16, warmup = argc > 3 ? atoi(argv[3]) : 3, size = argc > 4 ? atoi(argv[4]) : IMAGE_SIZE;
  cutracer::CudaRenderer* r = new cutracer::CudaRenderer();
  r->allocOutputImage(size, size);
  r->loadScene(argv[1]);
  r->setup();
  for (int i = 0; i < warmup; ++i) r->render();
  cudaDeviceSynchronize();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < frames; ++i) r->render();
  cudaDeviceSynchronize();
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const Image* img = r->getImage();
  double sum = 0; long finite = 0;
  for (long i = 0; i < 4L * img->width * img->height; ++i) { const float v = img->data[i]; if (v == v && v < 1e30f && v > -1e30f) { sum += v; ++finite; } }
  const double rays_max = 8.0 * size * size * SAMPLES_PER_PIXEL;
  fprintf(stderr, "\n");
  fflush(stdout);
  printf("\nREF_CUDA_JSON {\"scene\": \"%s\", \"size\": %d, \"spp_per_frame\": %d, \"frames\": %d, \"warmup\": %d, \"ms_per_frame\": %.4f, "
         "\"rays_per_frame_upper_bound\": %.0f, \"mrays_s_upper_bound\": %.2f, \"image_mean\": %.6f, \"finite_values\": %ld, \"cuda_error\": \"%s\"}\n",
         argv[1], size, SAMPLES_PER_PIXEL, frames, warmup, 1e3 * s / frames, rays_max, rays_max / (s / frames) / 1e6,
         finite ? sum / finite : 0.0, finite, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
