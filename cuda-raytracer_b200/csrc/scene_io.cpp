// Host helpers of the C ABI that need no device: .b2s scene files and camera placement.
// (b2rt_load_dae lives in dae_loader.cpp.)
#include <algorithm>
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
  if (!ok || hdr[0] != 1) { fclose(f); set_error(std::string(path) + ": not a version-1 .b2s scene"); return B2RT_ERR_IO; }
  const uint32_t nt = hdr[1], ns = hdr[2], nm = hdr[3], nl = hdr[4];
  // the header counts are checked against the file size before anything is allocated (a corrupt file must come back as
  // B2RT_ERR_IO, not as std::bad_alloc through an extern "C" function)
  const long body_at = ftell(f);
  long file_end = -1;
  if (body_at >= 0 && fseek(f, 0, SEEK_END) == 0) file_end = ftell(f);
  const uint64_t need = (uint64_t)nt * (72 + 4) + (uint64_t)ns * (16 + 4) + (uint64_t)nm * sizeof(b2rt_material) + (uint64_t)nl * sizeof(b2rt_light);
  if (body_at < 0 || file_end < 0 || fseek(f, body_at, SEEK_SET) != 0 || (uint64_t)(file_end - body_at) < need) {
    fclose(f); set_error(std::string(path) + ": truncated scene file (header counts exceed the file size)"); return B2RT_ERR_IO;
  }
  std::vector<float> tv, tn, sp;
  std::vector<uint32_t> tm, sm;
  std::vector<b2rt_material> mats;
  std::vector<b2rt_light> lights;
  try {
    tv.resize((size_t)nt * 9); tn.resize((size_t)nt * 9); sp.resize((size_t)ns * 4);
    tm.resize(nt); sm.resize(ns); mats.resize(nm); lights.resize(nl);
  } catch (const std::exception&) {
    fclose(f); set_error(std::string(path) + ": out of memory"); return B2RT_ERR_OOM;
  }
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
  bool ok = true;
  auto put = [&](const void* p, size_t size, size_t n) { if (n && fwrite(p, size, n, f) != n) ok = false; };
  put("B2S1", 1, 4); put(hdr, 4, 5); put(fh, 4, 11);
  std::vector<float> zeros;
  put(d.tri_verts, 4, (size_t)d.n_tris * 9);
  if (d.tri_normals) put(d.tri_normals, 4, (size_t)d.n_tris * 9);
  else { zeros.assign((size_t)d.n_tris * 9, 0.f); put(zeros.data(), 4, zeros.size()); }
  if (d.tri_material) put(d.tri_material, 4, d.n_tris);
  else { std::vector<uint32_t> z(d.n_tris, 0); put(z.data(), 4, z.size()); }
  put(d.spheres, 4, (size_t)d.n_spheres * 4);
  if (d.sphere_material) put(d.sphere_material, 4, d.n_spheres);
  else { std::vector<uint32_t> z(d.n_spheres, 0); put(z.data(), 4, z.size()); }
  put(d.materials, sizeof(b2rt_material), d.n_materials);
  put(d.lights, sizeof(b2rt_light), d.n_lights);
  if (fclose(f) != 0) ok = false;
  if (!ok) { set_error(std::string(path) + ": write failed"); return B2RT_ERR_IO; }
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

// ---- b2rt_camera_look_at: CudaRenderer::setViewpoint (src/cudaRenderer.cu:1845-1870, basis :1592-1599) -----------
extern "C" int b2rt_camera_look_at(const float origin[3], const float look_at[3], float fov_deg, b2rt_camera* out) {
  if (!origin || !look_at || !out) { b2rt::set_error("null argument"); return B2RT_ERR_INVALID; }
  auto cross = [](const double a[3], const double b[3], double r[3]) {
    r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
  };
  auto unit = [](double v[3]) { const double l = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); if (!(l > 0)) return false; v[0] /= l; v[1] /= l; v[2] /= l; return true; };
  double L[3] = {look_at[0], look_at[1], look_at[2]};
  if (!unit(L)) { b2rt::set_error("look_at is zero"); return B2RT_ERR_INVALID; }
  const double cdir[3] = {-L[0], -L[1], -L[2]}, acup[3] = {0, 1, 0};
  double left[3], up[3];
  cross(acup, cdir, left);
  if (!unit(left)) { b2rt::set_error("look_at is parallel to the y axis"); return B2RT_ERR_INVALID; }
  cross(left, cdir, up);
  unit(up);
  for (int a = 0; a < 3; ++a) {
    out->pos[a] = origin[a];
    out->c2w[a] = (float)left[a];        // screen right
    out->c2w[3 + a] = (float)-up[a];     // screen up
    out->c2w[6 + a] = (float)-L[a];      // towards the camera
  }
  const float fov = fov_deg > 0.f ? fov_deg : (float)(2.0 * std::atan(0.5) * 180.0 / 3.14159265358979323846);
  out->hfov_deg = out->vfov_deg = fov;
  return B2RT_OK;
}

// ---- image files -----------------------------------------------------------------------------------------------
namespace {
uint32_t crc32_png(const uint8_t* p, size_t n, uint32_t c) {
  static uint32_t T[256]; static bool init = false;
  if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t v = i; for (int k = 0; k < 8; ++k) v = (v & 1) ? 0xEDB88320u ^ (v >> 1) : v >> 1; T[i] = v; } init = true; }
  c = ~c; for (size_t i = 0; i < n; ++i) c = T[(c ^ p[i]) & 255] ^ (c >> 8); return ~c;
}
}  // namespace

// PNG, 8-bit RGBA, zlib stream of stored (uncompressed) deflate blocks; rows top to bottom = the buffer's rows last to first
extern "C" int b2rt_save_png(const char* path, const uint32_t* rgba8, uint32_t w, uint32_t h) {
  if (!path || !rgba8 || !w || !h) { b2rt::set_error("bad argument"); return B2RT_ERR_INVALID; }
  std::vector<uint8_t> raw;
  raw.reserve(((size_t)w * 4 + 1) * h);
  for (uint32_t y = 0; y < h; ++y) {
    raw.push_back(0);   // filter type none
    const uint8_t* row = reinterpret_cast<const uint8_t*>(rgba8 + (size_t)(h - 1 - y) * w);
    raw.insert(raw.end(), row, row + (size_t)w * 4);
  }
  FILE* f = fopen(path, "wb");
  if (!f) { b2rt::set_error(std::string("cannot write ") + path); return B2RT_ERR_IO; }
  bool ok = true;
  auto put = [&](const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; };
  auto be32 = [](uint32_t v, uint8_t* o) { o[0] = (uint8_t)(v >> 24); o[1] = (uint8_t)(v >> 16); o[2] = (uint8_t)(v >> 8); o[3] = (uint8_t)v; };
  auto chunk = [&](const char* tag, const std::vector<uint8_t>& data) {
    uint8_t len[4]; be32((uint32_t)data.size(), len); put(len, 4);
    std::vector<uint8_t> td(tag, tag + 4); td.insert(td.end(), data.begin(), data.end());
    put(td.data(), td.size());
    uint8_t c[4]; be32(crc32_png(td.data(), td.size(), 0), c); put(c, 4);
  };
  const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
  put(sig, 8);
  std::vector<uint8_t> ihdr(13); be32(w, &ihdr[0]); be32(h, &ihdr[4]); ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = ihdr[11] = ihdr[12] = 0;
  chunk("IHDR", ihdr);
  std::vector<uint8_t> z; z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;
  for (uint8_t v : raw) { a = (a + v) % 65521; b = (b + a) % 65521; }
  for (size_t pos = 0; pos < raw.size();) {
    const size_t n = std::min<size_t>(65535, raw.size() - pos);
    z.push_back(pos + n >= raw.size() ? 1 : 0);
    z.push_back((uint8_t)(n & 255)); z.push_back((uint8_t)(n >> 8)); z.push_back((uint8_t)(~n & 255)); z.push_back((uint8_t)((~n >> 8) & 255));
    z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
    pos += n;
  }
  uint8_t ad[4]; be32((b << 16) | a, ad); z.insert(z.end(), ad, ad + 4);
  chunk("IDAT", z);
  chunk("IEND", {});
  if (fclose(f) != 0) ok = false;
  if (!ok) { b2rt::set_error(std::string(path) + ": write failed"); return B2RT_ERR_IO; }
  return B2RT_OK;
}

// OpenEXR 2, single-part scanline image, NO_COMPRESSION, three FLOAT channels (stored B, G, R: the channel list is
// sorted by name), one scanline per block, increasing y = top row first
extern "C" int b2rt_save_exr(const char* path, const float* rgb, uint32_t w, uint32_t h) {
  if (!path || !rgb || !w || !h || w > 0x3FFFFFFu || h > 0x3FFFFFFu) { b2rt::set_error("bad argument"); return B2RT_ERR_INVALID; }
  std::vector<uint8_t> hd;
  auto p8 = [&](uint8_t v) { hd.push_back(v); };
  auto p32 = [&](uint32_t v) { for (int k = 0; k < 4; ++k) hd.push_back((uint8_t)(v >> (8 * k))); };
  auto pf = [&](float v) { uint32_t u; memcpy(&u, &v, 4); p32(u); };
  auto pstr = [&](const char* s) { while (*s) hd.push_back((uint8_t)*s++); hd.push_back(0); };
  auto attr = [&](const char* name, const char* type, uint32_t size) { pstr(name); pstr(type); p32(size); };
  p32(20000630u); p32(2u);   // magic, version 2 / no flags
  attr("channels", "chlist", 3 * 18 + 1);
  for (const char* c : {"B", "G", "R"}) { pstr(c); p32(2u /* FLOAT */); p8(0); p8(0); p8(0); p8(0); p32(1u); p32(1u); }
  p8(0);
  attr("compression", "compression", 1); p8(0);
  attr("dataWindow", "box2i", 16); p32(0); p32(0); p32(w - 1); p32(h - 1);
  attr("displayWindow", "box2i", 16); p32(0); p32(0); p32(w - 1); p32(h - 1);
  attr("lineOrder", "lineOrder", 1); p8(0);
  attr("pixelAspectRatio", "float", 4); pf(1.0f);
  attr("screenWindowCenter", "v2f", 8); pf(0.f); pf(0.f);
  attr("screenWindowWidth", "float", 4); pf(1.0f);
  p8(0);   // end of header
  const uint64_t line_bytes = (uint64_t)w * 12, block = 8 + line_bytes;
  uint64_t at = hd.size() + (uint64_t)h * 8;
  for (uint32_t y = 0; y < h; ++y) { for (int k = 0; k < 8; ++k) hd.push_back((uint8_t)(at >> (8 * k))); at += block; }
  FILE* f = fopen(path, "wb");
  if (!f) { b2rt::set_error(std::string("cannot write ") + path); return B2RT_ERR_IO; }
  bool ok = fwrite(hd.data(), 1, hd.size(), f) == hd.size();
  std::vector<float> line((size_t)w * 3);
  for (uint32_t y = 0; y < h && ok; ++y) {
    const float* row = rgb + (size_t)(h - 1 - y) * w * 3;
    for (uint32_t x = 0; x < w; ++x) { line[x] = row[3 * x + 2]; line[w + x] = row[3 * x + 1]; line[2 * (size_t)w + x] = row[3 * x]; }
    const uint32_t head[2] = {y, (uint32_t)line_bytes};
    ok = fwrite(head, 4, 2, f) == 2 && fwrite(line.data(), 4, line.size(), f) == line.size();
  }
  if (fclose(f) != 0) ok = false;
  if (!ok) { b2rt::set_error(std::string(path) + ": write failed"); return B2RT_ERR_IO; }
  return B2RT_OK;
}
