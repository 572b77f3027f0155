// Host helpers of the C ABI that need no device: .b2s scene files and camera placement.
// (b2rt_load_dae lives in dae_loader.cpp.)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "b2rt_internal.h"

using namespace b2rt;

namespace {
struct Storage {
  std::vector<float> tri_verts, tri_normals, spheres;
  std::vector<uint32_t> tri_material, sphere_material;
  std::vector<b2rt_material> materials;
  std::vector<b2rt_light> lights;
};
}  // namespace

namespace b2rt {
b2rt_scene_file* scene_file_from_parts(std::vector<float>&& tv, std::vector<float>&& tn, std::vector<uint32_t>&& tm,
                                       std::vector<float>&& sp, std::vector<uint32_t>&& sm, std::vector<b2rt_material>&& mats,
                                       std::vector<b2rt_light>&& lights, const float cam_dir[3], float hfov, float vfov,
                                       const float bbox[6]) {
  Storage* st = new Storage();
  st->tri_verts = std::move(tv); st->tri_normals = std::move(tn); st->tri_material = std::move(tm);
  st->spheres = std::move(sp); st->sphere_material = std::move(sm); st->materials = std::move(mats); st->lights = std::move(lights);
  b2rt_scene_file* f = new b2rt_scene_file();
  memset(f, 0, sizeof *f);
  f->storage = st;
  f->desc.n_tris = (uint32_t)(st->tri_verts.size() / 9);
  f->desc.tri_verts = st->tri_verts.data();
  f->desc.tri_normals = st->tri_normals.size() == st->tri_verts.size() && !st->tri_normals.empty() ? st->tri_normals.data() : nullptr;
  f->desc.tri_material = st->tri_material.data();
  f->desc.n_spheres = (uint32_t)(st->spheres.size() / 4);
  f->desc.spheres = st->spheres.data();
  f->desc.sphere_material = st->sphere_material.data();
  f->desc.n_materials = (uint32_t)st->materials.size(); f->desc.materials = st->materials.data();
  f->desc.n_lights = (uint32_t)st->lights.size(); f->desc.lights = st->lights.data();
  memcpy(f->cam_dir, cam_dir, 12); f->cam_hfov_deg = hfov; f->cam_vfov_deg = vfov; memcpy(f->bbox, bbox, 24);
  b2rt_camera_place(f->bbox, f->cam_dir, hfov, vfov, 4, 3, &f->camera);
  return f;
}
}  // namespace b2rt

extern "C" {

int b2rt_scene_load(const char* path, b2rt_scene_file** out) {
  if (!path || !out) { set_error("null argument"); return B2RT_ERR_INVALID; }
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { set_error(std::string("cannot open ") + path); return B2RT_ERR_IO; }
  char magic[4]; uint32_t hdr[5]; float fh[11];
  bool ok = fread(magic, 1, 4, f) == 4 && memcmp(magic, "B2S1", 4) == 0 && fread(hdr, 4, 5, f) == 5 && fread(fh, 4, 11, f) == 11;
  if (!ok) { fclose(f); set_error(std::string(path) + ": not a .b2s scene"); return B2RT_ERR_IO; }
  const uint32_t nt = hdr[1], ns = hdr[2], nm = hdr[3], nl = hdr[4];
  std::vector<float> tv((size_t)nt * 9), tn((size_t)nt * 9), sp((size_t)ns * 4);
  std::vector<uint32_t> tm(nt), sm(ns);
  std::vector<b2rt_material> mats(nm);
  std::vector<b2rt_light> lights(nl);
  ok = fread(tv.data(), 4, tv.size(), f) == tv.size() && fread(tn.data(), 4, tn.size(), f) == tn.size() &&
       fread(tm.data(), 4, nt, f) == nt && fread(sp.data(), 4, sp.size(), f) == sp.size() && fread(sm.data(), 4, ns, f) == ns &&
       fread(mats.data(), sizeof(b2rt_material), nm, f) == nm && fread(lights.data(), sizeof(b2rt_light), nl, f) == nl;
  fclose(f);
  if (!ok) { set_error(std::string(path) + ": truncated scene file"); return B2RT_ERR_IO; }
  *out = scene_file_from_parts(std::move(tv), std::move(tn), std::move(tm), std::move(sp), std::move(sm), std::move(mats),
                               std::move(lights), fh, fh[3], fh[4], fh + 5);
  return B2RT_OK;
}

int b2rt_scene_save(const char* path, const b2rt_scene_file* s) {
  if (!path || !s) { set_error("null argument"); return B2RT_ERR_INVALID; }
  FILE* f = fopen(path, "wb");
  if (!f) { set_error(std::string("cannot open ") + path); return B2RT_ERR_IO; }
  const b2rt_scene_desc& d = s->desc;
  uint32_t hdr[5] = {1, d.n_tris, d.n_spheres, d.n_materials, d.n_lights};
  float fh[11] = {s->cam_dir[0], s->cam_dir[1], s->cam_dir[2], s->cam_hfov_deg, s->cam_vfov_deg,
                  s->bbox[0], s->bbox[1], s->bbox[2], s->bbox[3], s->bbox[4], s->bbox[5]};
  fwrite("B2S1", 1, 4, f); fwrite(hdr, 4, 5, f); fwrite(fh, 4, 11, f);
  std::vector<float> zeros;
  fwrite(d.tri_verts, 4, (size_t)d.n_tris * 9, f);
  if (d.tri_normals) fwrite(d.tri_normals, 4, (size_t)d.n_tris * 9, f);
  else { zeros.assign((size_t)d.n_tris * 9, 0.f); fwrite(zeros.data(), 4, zeros.size(), f); }
  if (d.tri_material) fwrite(d.tri_material, 4, d.n_tris, f);
  else { std::vector<uint32_t> z(d.n_tris, 0); fwrite(z.data(), 4, z.size(), f); }
  fwrite(d.spheres, 4, (size_t)d.n_spheres * 4, f);
  if (d.sphere_material) fwrite(d.sphere_material, 4, d.n_spheres, f);
  else { std::vector<uint32_t> z(d.n_spheres, 0); fwrite(z.data(), 4, z.size(), f); }
  fwrite(d.materials, sizeof(b2rt_material), d.n_materials, f);
  fwrite(d.lights, sizeof(b2rt_light), d.n_lights, f);
  fclose(f);
  return B2RT_OK;
}

void b2rt_scene_free(b2rt_scene_file* s) {
  if (!s) return;
  delete static_cast<Storage*>(s->storage);
  delete s;
}

// Application::load camera placement (src/application.cpp:395-408) + Camera::place / compute_position
// (src/camera.cpp:35-50, 87-109) + Camera::configure fov fix-up (src/camera.cpp:15-33).
int b2rt_camera_place(const float bbox[6], const float view_dir[3], float hfov_deg, float vfov_deg, uint32_t width,
                      uint32_t height, b2rt_camera* out) {
  if (!bbox || !view_dir || !out || width == 0 || height == 0) { set_error("bad argument"); return B2RT_ERR_INVALID; }
  const double PI = 3.14159265358979323846;
  double target[3], ext[3];
  for (int a = 0; a < 3; ++a) { target[a] = 0.5 * ((double)bbox[a] + (double)bbox[3 + a]); ext[a] = (double)bbox[3 + a] - (double)bbox[a]; }
  const double canonical = std::sqrt(ext[0] * ext[0] + ext[1] * ext[1] + ext[2] * ext[2]) / 2 * 1.5;
  double r = canonical * 2;
  r = std::min(std::max(r, canonical / 10.0), canonical * 20.0);
  double dl = std::sqrt((double)view_dir[0] * view_dir[0] + (double)view_dir[1] * view_dir[1] + (double)view_dir[2] * view_dir[2]);
  if (!(dl > 0)) { set_error("zero view direction"); return B2RT_ERR_INVALID; }
  const double cd[3] = {view_dir[0] / dl, view_dir[1] / dl, view_dir[2] / dl};
  double phi = std::acos(std::max(-1.0, std::min(1.0, cd[1])));
  const double theta = std::atan2(cd[0], cd[2]);
  if (std::sin(phi) == 0) phi += 1e-5;
  const double sp = std::sin(phi);
  const double tc[3] = {r * sp * std::sin(theta), r * std::cos(phi), r * sp * std::cos(theta)};
  const double up[3] = {0, sp > 0 ? 1.0 : -1.0, 0};
  auto cross = [](const double* a, const double* b, double* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
  auto norm = [](double* v) { double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); v[0] /= l; v[1] /= l; v[2] /= l; };
  double x[3], y[3], z[3] = {tc[0], tc[1], tc[2]};
  cross(up, tc, x); norm(x);
  cross(tc, x, y); norm(y);
  norm(z);
  double hf = hfov_deg, vf = vfov_deg;
  const double ar1 = std::tan(hf * PI / 360) / std::tan(vf * PI / 360), ar = (double)width / height;
  if (ar1 < ar) hf = 2 * std::atan(std::tan(vf * PI / 360) * ar) * 180 / PI;
  else if (ar1 > ar) vf = 2 * std::atan(std::tan(hf * PI / 360) / ar) * 180 / PI;
  for (int a = 0; a < 3; ++a) {
    out->pos[a] = (float)(target[a] + tc[a]);
    out->c2w[a] = (float)x[a]; out->c2w[3 + a] = (float)y[a]; out->c2w[6 + a] = (float)z[a];
  }
  out->hfov_deg = (float)hf; out->vfov_deg = (float)vf;
  return B2RT_OK;
}

int b2rt_load_dae(const char* path, b2rt_scene_file** out);

}  // extern "C"
