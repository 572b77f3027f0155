// ORACLE / TEST INFRASTRUCTURE -- second reported baseline: the reference's OWN CUDA renderer
// (src/cudaRenderer.cu, compiled unmodified for sm_100a from where it lies under /root/reference) driven
// headless.  This file is ours: it replaces src/cudaMain.cpp + src/display.cpp (GLUT window) with the same
// call sequence (cudaMain.cpp:88-99: allocOutputImage -> loadScene -> setup, then display.cpp:120-136:
// render() per displayed frame) and times the frames with a host clock around cudaDeviceSynchronize.
//
//   ref_cuda_render <scene.dae> [frames=16] [warmup=3] [size=512]
//   ref_cuda_render --dump-primary <out.bin> <scene.dae> [size=512]
//
// --dump-primary (parity pin, SURVEY 8c: "cross-check against the rebuilt reference kernel compares t per ray id"): renders
// ONE frame and writes (a) the camera rays the reference generated and (b) what its own traversal found for them --
// CuIntersection.t / .valid of deviceIntersections after the primary rayIntersect() (src/cudaRenderer.cu:2304-2331,
// merge :515-540) -- plus (c) the triangles its loader produced (the host vector CudaRenderer::triangles,
// src/cudaRenderer.cu:1760-1792).  The reference sources stay unmodified: the class is read through `#define private
// public` in THIS translation unit only, and the two device buffers are copied out from a std::cout stream-buffer hook
// that fires when CudaRenderer::lapTimer prints the stage names "PrimaryRays()" and "Primary Ray Intersect"
// (src/cudaRenderer.cu:2499-2513; the device is synchronised at both points).
//
// The reference renders SAMPLES_PER_PIXEL (2) samples per render() on a square power-of-two image with a
// fixed script of 8 traversal passes (primary, 2 + 2 + 1 shadow passes, 2 scene bounces:
// cudaRenderer.cu:2499-2534), so a frame traces at most 8 * size^2 * 2 rays.  Prints one JSON line.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <streambuf>
#include <string>
#include <vector>

#define private public      // this translation unit only: read CudaRenderer's device pointers (layout is unaffected)
#include "cudaRenderer.h"
#undef private

// the reference's error dialog (src/error_dialog.cpp) needs GLUT; collada.cpp only calls this entry point
void showError(std::string msg, bool fatal) {
  fprintf(stderr, "reference error: %s\n", msg.c_str());
  if (fatal) exit(1);
}

// ---- --dump-primary ---------------------------------------------------------------------------------------------
namespace {
struct DumpHook : std::streambuf {
  cutracer::CudaRenderer* r = nullptr;
  size_t n = 0;
  std::vector<cutracer::CuRay> rays;
  std::vector<cutracer::CuIntersection> hits;
  std::string tail;
  bool got_rays = false, got_hits = false;
  void feed(const char* s, std::streamsize k) {
    tail.append(s, (size_t)k);
    if (!got_rays && tail.find("PrimaryRays()") != std::string::npos) {
      rays.resize(n);
      cudaMemcpy(rays.data(), r->deviceRays1, n * sizeof(cutracer::CuRay), cudaMemcpyDeviceToHost);
      got_rays = true;
    }
    if (!got_hits && tail.find("Primary Ray Intersect") != std::string::npos) {
      hits.resize(n);
      cudaMemcpy(hits.data(), r->deviceIntersections, n * sizeof(cutracer::CuIntersection), cudaMemcpyDeviceToHost);
      got_hits = true;
    }
    if (tail.size() > 256) tail.erase(0, tail.size() - 64);
  }
  std::streamsize xsputn(const char* s, std::streamsize k) override { feed(s, k); return k; }
  int overflow(int c) override { if (c != EOF) { char ch = (char)c; feed(&ch, 1); } return c; }
};

int dump_primary(const char* out_path, const char* scene, int size) {
  cutracer::CudaRenderer* r = new cutracer::CudaRenderer();
  r->allocOutputImage(size, size);
  r->loadScene(scene);
  r->setup();
  DumpHook hook;
  hook.r = r; hook.n = (size_t)size * size * SAMPLES_PER_PIXEL;
  std::streambuf* old = std::cout.rdbuf(&hook);
  r->render();
  cudaDeviceSynchronize();
  std::cout.rdbuf(old);
  if (!hook.got_rays || !hook.got_hits) { fprintf(stderr, "dump: stage markers not seen\n"); return 1; }
  FILE* f = fopen(out_path, "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", out_path); return 1; }
  const unsigned n = (unsigned)hook.n, nt = (unsigned)r->triangles.size();
  const unsigned hdr[4] = {0x31504652u /* "RFP1" */, n, nt, (unsigned)size};
  fwrite(hdr, 4, 4, f);
  for (unsigned i = 0; i < n; ++i) {       // the ray with id i, and the reference's merged intersection for id i
    const cutracer::CuRay& q = hook.rays[i];
    const cutracer::CuIntersection& h = hook.hits[i];
    const float rec[8] = {q.o.x, q.o.y, q.o.z, q.d.x, q.d.y, q.d.z, h.t, h.valid ? 1.f : 0.f};
    fwrite(rec, 4, 8, f);
  }
  for (unsigned i = 0; i < nt; ++i) {      // loader output: positions + shading normals + material class per triangle
    const cutracer::CuTriangle& t = r->triangles[i];
    const cutracer::CuBSDF& b = r->bsdfs[t.bsdf];
    const float rec[24] = {t.a.x, t.a.y, t.a.z, t.b.x, t.b.y, t.b.z, t.c.x, t.c.y, t.c.z, t.n0.x, t.n0.y, t.n0.z, t.n1.x, t.n1.y, t.n1.z,
                           t.n2.x, t.n2.y, t.n2.z, (float)b.fn, b.albedo.x, b.albedo.y, b.albedo.z, b.radiance.x, (float)t.emit};
    fwrite(rec, 4, 24, f);
  }
  fclose(f);
  const cudaError_t e = cudaGetLastError();
  fprintf(stderr, "dump: %u rays, %u triangles -> %s (%s)\n", n, nt, out_path, cudaGetErrorString(e));
  return 0;
}
}  // namespace

int main(int argc, char** argv) {
  if (argc >= 4 && !strcmp(argv[1], "--dump-primary")) return dump_primary(argv[2], argv[3], argc > 4 ? atoi(argv[4]) : IMAGE_SIZE);
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
