// b2rt_load_dae: COLLADA (.dae, CMU462 profile) loader + mesh flattening, host C++.
//
// Replaces, for the path-tracing hot path only:
//   Collada::ColladaParser::load / parse_node / parse_polymesh / parse_material / parse_light / parse_sphere /
//   parse_camera                                   src/collada/collada.cpp:117-951
//     up_axis fix-up matrix                        :151-190
//     node <matrix> wins ('break' after it)        :232-256
//     material: CMU462 <extra> technique wins, else phong/diffuse  :864-951
//   Application::load camera-direction rule c_dir = unit(M * (view_dir, 1))   src/application.cpp:366-367
//   sphere centre / scale                          src/application.cpp:474-479
//   DynamicScene::Mesh (vertices through the node matrix)          src/dynamic_scene/mesh.cpp:21-46
//   StaticScene::Mesh (first three vertices of a polygon; HalfedgeMesh::triangulate is a stub)
//                                                  src/static_scene/object.cpp:17-72, src/meshEdit.cpp:360-364
//   Vertex::normal() area-weighted normal          src/halfEdgeMesh.h:619-644
//   DynamicScene::AreaLight frame from the node matrix             src/dynamic_scene/area_light.h:12-24
// Error behaviour: the reference returns < 0 when the file cannot be opened and exit()s on malformed XML
// (collada.cpp:117-214); here every failure is B2RT_ERR_IO / B2RT_ERR_INVALID with b2rt_last_error() set.
//
// The XML reader below is a small non-validating DOM parser (elements, attributes, character data, comments,
// processing instructions, CDATA, the five predefined entities) -- enough for COLLADA 1.4 exporters; the
// reference links tinyxml2 for the same job.  All geometry is computed in double and narrowed to float at the end,
// the same arithmetic as tools/dae2scene.py (the offline converter this loader supersedes).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "b2rt_internal.h"

namespace b2rt {
b2rt_scene_file* scene_file_from_parts(std::vector<float>&& tv, std::vector<float>&& tn, std::vector<uint32_t>&& tm,
                                       std::vector<float>&& sp, std::vector<uint32_t>&& sm, std::vector<b2rt_material>&& mats,
                                       std::vector<b2rt_light>&& lights, const float cam_dir[3], float hfov, float vfov,
                                       const float bbox[6]);
}

namespace {

// ---- XML ------------------------------------------------------------------------------------------------
struct XmlNode {
  std::string tag;                                   // local name (namespace prefix stripped)
  std::vector<std::pair<std::string, std::string>> attrs;
  std::string text;                                  // character data before the first child element
  std::vector<std::unique_ptr<XmlNode>> kids;
  const std::string* attr(const char* name) const {
    for (auto& a : attrs) if (a.first == name) return &a.second;
    return nullptr;
  }
  const XmlNode* child(const char* name) const {
    for (auto& k : kids) if (k->tag == name) return k.get();
    return nullptr;
  }
  // direct-child path "a/b/c"
  const XmlNode* find(const char* path) const {
    const XmlNode* n = this;
    std::string p(path);
    size_t s = 0;
    while (n && s <= p.size()) {
      size_t e = p.find('/', s);
      if (e == std::string::npos) e = p.size();
      n = n->child(p.substr(s, e - s).c_str());
      s = e + 1;
      if (e == p.size()) break;
    }
    return n;
  }
};

struct XmlError { std::string msg; };

class XmlParser {
 public:
  XmlParser(const char* b, size_t n) : p_(b), end_(b + n), begin_(b) {}
  std::unique_ptr<XmlNode> parse_document() {
    skip_misc();
    if (p_ >= end_ || *p_ != '<') fail("no root element");
    auto root = parse_element(0);
    skip_misc();
    return root;
  }

 private:
  const char* p_; const char* end_; const char* begin_;
  [[noreturn]] void fail(const std::string& what) {
    size_t line = 1;
    for (const char* q = begin_; q < p_ && q < end_; ++q) line += *q == '\n';
    throw XmlError{"XML error at line " + std::to_string(line) + ": " + what};
  }
  static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
  static bool is_name(char c) { return isalnum((unsigned char)c) || c == '_' || c == '-' || c == '.' || c == ':'; }
  bool starts(const char* s) const { size_t n = strlen(s); return (size_t)(end_ - p_) >= n && memcmp(p_, s, n) == 0; }
  void skip_until(const char* s) {
    size_t n = strlen(s);
    while ((size_t)(end_ - p_) >= n && memcmp(p_, s, n) != 0) ++p_;
    if ((size_t)(end_ - p_) < n) fail(std::string("unterminated construct, expected ") + s);
    p_ += n;
  }
  void skip_misc() {   // whitespace, BOM, comments, processing instructions, DOCTYPE
    for (;;) {
      while (p_ < end_ && (is_space(*p_) || (unsigned char)*p_ >= 0xEF)) ++p_;
      if (starts("<!--")) { p_ += 4; skip_until("-->"); }
      else if (starts("<?")) { p_ += 2; skip_until("?>"); }
      else if (starts("<!DOCTYPE")) { p_ += 9; skip_until(">"); }
      else return;
    }
  }
  static void append_entity(const std::string& e, std::string* out) {
    if (e == "lt") *out += '<';
    else if (e == "gt") *out += '>';
    else if (e == "amp") *out += '&';
    else if (e == "quot") *out += '"';
    else if (e == "apos") *out += '\'';
    else if (!e.empty() && e[0] == '#') {
      long v = e.size() > 1 && (e[1] == 'x' || e[1] == 'X') ? strtol(e.c_str() + 2, nullptr, 16) : strtol(e.c_str() + 1, nullptr, 10);
      if (v > 0 && v < 128) *out += (char)v; else *out += '?';
    } else { *out += '&'; *out += e; *out += ';'; }
  }
  std::string decode(const char* b, const char* e) {
    std::string out;
    out.reserve(e - b);
    while (b < e) {
      if (*b == '&') {
        const char* s = (const char*)memchr(b, ';', e - b);
        if (!s) { out.append(b, e); break; }
        append_entity(std::string(b + 1, s), &out);
        b = s + 1;
      } else out += *b++;
    }
    return out;
  }
  std::string parse_name() {
    const char* s = p_;
    while (p_ < end_ && is_name(*p_)) ++p_;
    if (p_ == s) fail("expected a name");
    std::string n(s, p_);
    size_t c = n.rfind(':');
    return c == std::string::npos ? n : n.substr(c + 1);
  }
  std::unique_ptr<XmlNode> parse_element(int depth) {
    if (depth > 256) fail("element nesting too deep");
    ++p_;   // '<'
    auto n = std::make_unique<XmlNode>();
    n->tag = parse_name();
    for (;;) {   // attributes
      while (p_ < end_ && is_space(*p_)) ++p_;
      if (p_ >= end_) fail("unterminated start tag <" + n->tag);
      if (*p_ == '/') { if (!starts("/>")) fail("malformed empty-element tag"); p_ += 2; return n; }
      if (*p_ == '>') { ++p_; break; }
      const char* s = p_;
      while (p_ < end_ && is_name(*p_)) ++p_;
      if (p_ == s) fail("malformed attribute in <" + n->tag + ">");
      std::string an(s, p_);
      while (p_ < end_ && is_space(*p_)) ++p_;
      if (p_ >= end_ || *p_ != '=') fail("attribute without value in <" + n->tag + ">");
      ++p_;
      while (p_ < end_ && is_space(*p_)) ++p_;
      if (p_ >= end_ || (*p_ != '"' && *p_ != '\'')) fail("unquoted attribute value in <" + n->tag + ">");
      const char q = *p_++;
      const char* vs = p_;
      while (p_ < end_ && *p_ != q) ++p_;
      if (p_ >= end_) fail("unterminated attribute value");
      n->attrs.emplace_back(an, decode(vs, p_));
      ++p_;
    }
    bool seen_child = false;
    for (;;) {   // content
      const char* s = p_;
      while (p_ < end_ && *p_ != '<') ++p_;
      if (p_ >= end_) fail("unterminated element <" + n->tag + ">");
      if (!seen_child && p_ > s) n->text += decode(s, p_);
      if (starts("<!--")) { p_ += 4; skip_until("-->"); }
      else if (starts("<![CDATA[")) {
        p_ += 9; const char* cs = p_; skip_until("]]>");
        if (!seen_child) n->text.append(cs, p_ - 3);
      } else if (starts("<?")) { p_ += 2; skip_until("?>"); }
      else if (starts("</")) {
        p_ += 2;
        std::string cn = parse_name();
        if (cn != n->tag) fail("mismatched end tag </" + cn + "> for <" + n->tag + ">");
        while (p_ < end_ && is_space(*p_)) ++p_;
        if (p_ >= end_ || *p_ != '>') fail("malformed end tag");
        ++p_;
        return n;
      } else {
        n->kids.push_back(parse_element(depth + 1));
        seen_child = true;
      }
    }
  }
};

// ---- small double-precision helpers ---------------------------------------------------------------------
struct M4 { double m[4][4]; };
M4 identity() { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = i == j; return r; }
M4 mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a.m[i][k] * b.m[k][j];
      r.m[i][j] = s;
    }
  return r;
}
struct V3 { double x, y, z; };
V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
double length(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
// M * (p, 1), first three components (no perspective divide: Matrix4x4 * Vector4D then .to3D(), like the converter)
V3 xf_point(const M4& M, V3 p) {
  return {M.m[0][0] * p.x + M.m[0][1] * p.y + M.m[0][2] * p.z + M.m[0][3],
          M.m[1][0] * p.x + M.m[1][1] * p.y + M.m[1][2] * p.z + M.m[1][3],
          M.m[2][0] * p.x + M.m[2][1] * p.y + M.m[2][2] * p.z + M.m[2][3]};
}

bool parse_doubles(const std::string& s, std::vector<double>* out) {
  const char* p = s.c_str();
  for (;;) {
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p;
    if (!*p) return true;
    char* e = nullptr;
    double v = strtod(p, &e);
    if (e == p) return false;
    out->push_back(v);
    p = e;
  }
}
bool parse_ints(const std::string& s, std::vector<long long>* out) {
  const char* p = s.c_str();
  for (;;) {
    while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') ++p;
    if (!*p) return true;
    char* e = nullptr;
    long long v = strtoll(p, &e, 10);
    if (e == p) return false;
    out->push_back(v);
    p = e;
  }
}

struct LoadError { std::string msg; };
[[noreturn]] void bad(const std::string& m) { throw LoadError{m}; }

// get_technique_common / get_technique_cmu462 (collada.cpp): first match in document order in the subtree
const XmlNode* find_first(const XmlNode* e, const char* tag, const char* attr = nullptr, const char* value = nullptr) {
  if (e->tag == tag) {
    if (!attr) return e;
    const std::string* a = e->attr(attr);
    if (a && *a == value) return e;
  }
  for (auto& k : e->kids)
    if (const XmlNode* r = find_first(k.get(), tag, attr, value)) return r;
  return nullptr;
}
const XmlNode* technique_common(const XmlNode* e) {
  if (const XmlNode* t = find_first(e, "technique_common")) return t;
  // every profile_COMMON in document order; the first one that has a <technique> child
  struct Walk {
    static const XmlNode* go(const XmlNode* n) {
      if (n->tag == "profile_COMMON") return n->child("technique");   // nullptr => the caller moves on to the next one
      for (auto& k : n->kids) if (const XmlNode* r = go(k.get())) return r;
      return nullptr;
    }
  };
  return Walk::go(e);
}
const XmlNode* technique_cmu462(const XmlNode* e) { return find_first(e, "technique", "profile", "CMU462"); }

struct Loader {
  std::map<std::string, const XmlNode*> ids;
  std::vector<float> tri_verts, tri_normals, spheres;
  std::vector<uint32_t> tri_material, sphere_material;
  std::vector<b2rt_material> materials;
  std::vector<b2rt_light> lights;
  std::map<std::string, uint32_t> mat_index;
  double cam_dir[3] = {0, 0, 1};
  double hfov = 50.0, vfov = 35.0;
  double bmin[3] = {INFINITY, INFINITY, INFINITY}, bmax[3] = {-INFINITY, -INFINITY, -INFINITY};

  const XmlNode* by_url(const XmlNode* e, const char* attr, const char* what) {
    const std::string* u = e->attr(attr);
    if (!u || u->empty()) bad(std::string(what) + ": missing '" + attr + "'");
    auto it = ids.find((*u)[0] == '#' ? u->substr(1) : *u);
    if (it == ids.end()) bad(std::string(what) + ": unresolved reference " + *u);
    return it->second;
  }
  static std::vector<double> floats_of(const XmlNode* e, const char* what, size_t at_least) {
    std::vector<double> v;
    if (!e || !parse_doubles(e->text, &v) || v.size() < at_least) bad(std::string("bad or missing <") + what + ">");
    return v;
  }
  void grow_bbox(V3 lo, V3 hi) {
    bmin[0] = std::fmin(bmin[0], lo.x); bmin[1] = std::fmin(bmin[1], lo.y); bmin[2] = std::fmin(bmin[2], lo.z);
    bmax[0] = std::fmax(bmax[0], hi.x); bmax[1] = std::fmax(bmax[1], hi.y); bmax[2] = std::fmax(bmax[2], hi.z);
  }

  // parse_material (collada.cpp:864-951)
  b2rt_material material(const std::string& id) {
    auto it = ids.find(id);
    if (it == ids.end()) bad("unresolved material #" + id);
    const XmlNode* inst = it->second->child("instance_effect");
    if (!inst) bad("material " + id + ": no <instance_effect>");
    const XmlNode* eff = by_url(inst, "url", "instance_effect");
    b2rt_material m;
    memset(&m, 0, sizeof m);
    m.kind = B2RT_MAT_DIFFUSE; m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.5f; m.ior = 1.0f;
    auto set3 = [](float* dst, const std::vector<double>& v) { for (int i = 0; i < 3; ++i) dst[i] = (float)v[i]; };
    auto zero3 = [](float* dst) { dst[0] = dst[1] = dst[2] = 0.f; };
    if (const XmlNode* t = technique_cmu462(eff)) {
      for (auto& bp : t->kids) {
        const XmlNode* b = bp.get();
        if (b->tag == "emission") {
          m.kind = B2RT_MAT_EMISSION; set3(m.emission, floats_of(b->child("radiance"), "radiance", 3)); zero3(m.albedo);
        } else if (b->tag == "mirror") {
          m.kind = B2RT_MAT_MIRROR; set3(m.albedo, floats_of(b->child("reflectance"), "reflectance", 3));
        } else if (b->tag == "glossy") {   // commented out in the reference's parser (collada.cpp:898-907); accepted here
          m.kind = B2RT_MAT_GLOSSY; set3(m.albedo, floats_of(b->child("reflectance"), "reflectance", 3));
          m.roughness = (float)floats_of(b->child("roughness"), "roughness", 1)[0];
        } else if (b->tag == "refraction") {
          m.kind = B2RT_MAT_REFRACTION; set3(m.transmittance, floats_of(b->child("transmittance"), "transmittance", 3));
          m.roughness = (float)floats_of(b->child("roughness"), "roughness", 1)[0];
          m.ior = (float)floats_of(b->child("ior"), "ior", 1)[0];
          zero3(m.albedo);
        } else if (b->tag == "glass") {
          m.kind = B2RT_MAT_GLASS; set3(m.transmittance, floats_of(b->child("transmittance"), "transmittance", 3));
          set3(m.albedo, floats_of(b->child("reflectance"), "reflectance", 3));
          m.roughness = (float)floats_of(b->child("roughness"), "roughness", 1)[0];
          m.ior = (float)floats_of(b->child("ior"), "ior", 1)[0];
        }
      }
    } else if (const XmlNode* t = technique_common(eff)) {
      if (const XmlNode* d = t->find("phong/diffuse/color")) set3(m.albedo, floats_of(d, "color", 3));
    }
    return m;
  }

  uint32_t material_for(const XmlNode* node) {
    const XmlNode* im = node->find("instance_geometry/bind_material/technique_common/instance_material");
    std::string key;
    if (!im) {
      key = "__default_white__";
      auto it = mat_index.find(key);
      if (it != mat_index.end()) return it->second;
      b2rt_material m;
      memset(&m, 0, sizeof m);
      m.kind = B2RT_MAT_DIFFUSE; m.albedo[0] = m.albedo[1] = m.albedo[2] = 1.f; m.ior = 1.f;
      mat_index[key] = (uint32_t)materials.size();
      materials.push_back(m);
      return mat_index[key];
    }
    const std::string* tgt = im->attr("target");
    if (!tgt || tgt->empty()) bad("instance_material without target");
    key = (*tgt)[0] == '#' ? tgt->substr(1) : *tgt;
    auto it = mat_index.find(key);
    if (it != mat_index.end()) return it->second;
    const uint32_t idx = (uint32_t)materials.size();
    mat_index[key] = idx;
    materials.push_back(material(key));
    return idx;
  }

  // parse_polymesh + Mesh flattening + vertex normals
  void add_mesh(const XmlNode* mesh, const M4& M, uint32_t mat) {
    std::map<std::string, std::vector<double>> sources;
    for (auto& s : mesh->kids) {
      if (s->tag != "source") continue;
      const XmlNode* fa = s->child("float_array");
      const std::string* id = s->attr("id");
      if (!fa || !id) continue;
      std::vector<double> v;
      if (!parse_doubles(fa->text, &v)) bad("mesh source " + *id + ": malformed float_array");
      sources[*id] = std::move(v);
    }
    const XmlNode* verts = mesh->child("vertices");
    if (!verts) bad("mesh without <vertices>");
    std::string pos_src;
    for (auto& inp : verts->kids) {
      if (inp->tag != "input") continue;
      const std::string* sem = inp->attr("semantic"); const std::string* src = inp->attr("source");
      if (sem && src && *sem == "POSITION") pos_src = (*src)[0] == '#' ? src->substr(1) : *src;
    }
    auto ps = sources.find(pos_src);
    if (ps == sources.end()) bad("mesh: POSITION source not found");
    const std::vector<double>& raw = ps->second;
    const size_t nv = raw.size() / 3;
    std::vector<V3> P(nv);
    V3 lo = {INFINITY, INFINITY, INFINITY}, hi = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i < nv; ++i) {
      const double x = raw[3 * i], y = raw[3 * i + 1], z = raw[3 * i + 2];
      double h[4];
      for (int r = 0; r < 4; ++r) h[r] = x * M.m[r][0] + y * M.m[r][1] + z * M.m[r][2] + 1.0 * M.m[r][3];
      P[i] = {h[0] / h[3], h[1] / h[3], h[2] / h[3]};   // projectTo3D of the homogeneous product (mesh.cpp:33-36)
      lo = {std::fmin(lo.x, P[i].x), std::fmin(lo.y, P[i].y), std::fmin(lo.z, P[i].z)};
      hi = {std::fmax(hi.x, P[i].x), std::fmax(hi.y, P[i].y), std::fmax(hi.z, P[i].z)};
    }
    const XmlNode* pl = mesh->child("polylist");
    const bool is_poly = pl != nullptr;
    if (!pl) pl = mesh->child("triangles");
    if (!pl) bad("mesh has neither <polylist> nor <triangles>");
    long long off_vertex = -1; int stride = 0;
    {
      bool hv = false, hn = false, ht = false;
      for (auto& inp : pl->kids) {
        if (inp->tag != "input") continue;
        const std::string* sem = inp->attr("semantic"); const std::string* off = inp->attr("offset");
        if (!sem || !off) continue;
        if (*sem == "VERTEX") { hv = true; off_vertex = atoll(off->c_str()); }
        else if (*sem == "NORMAL") hn = true;
        else if (*sem == "TEXCOORD") ht = true;
      }
      stride = (int)hv + (int)hn + (int)ht;   // the reference strides by the number of known semantics (collada.cpp:805-860)
      if (!hv) bad("polylist without a VERTEX input");
    }
    const std::string* cnt = pl->attr("count");
    if (!cnt) bad("polylist without count");
    const long long npoly = atoll(cnt->c_str());
    if (npoly < 0 || npoly > 0x7FFFFFFF) bad("polylist: bad polygon count");
    std::vector<long long> sizes;
    if (is_poly) {
      const XmlNode* vc = pl->child("vcount");
      if (!vc || !parse_ints(vc->text, &sizes) || (long long)sizes.size() < npoly) bad("polylist: bad <vcount>");
      sizes.resize(npoly);
    } else sizes.assign(npoly, 3);
    std::vector<long long> idx;
    const XmlNode* pe = pl->child("p");
    if (!pe || !parse_ints(pe->text, &idx)) bad("polylist: bad <p>");
    std::vector<uint32_t> tri((size_t)npoly * 3);
    long long start = 0;
    for (long long f = 0; f < npoly; ++f) {
      if (sizes[f] < 3) bad("polygon with fewer than three vertices");
      for (int k = 0; k < 3; ++k) {
        const long long at = (start + k) * stride + off_vertex;
        if (at < 0 || at >= (long long)idx.size()) bad("polylist: index list too short");
        const long long v = idx[at];
        if (v < 0 || v >= (long long)nv) bad("polylist: vertex index out of range");
        tri[3 * f + k] = (uint32_t)v;
      }
      start += sizes[f];
    }
    // Vertex::normal(): sum over incident faces of cross(pj - pi, pk - pi) == the face's un-normalised normal at
    // every corner of a triangle; accumulation order = corner-major (all first corners, then second, then third)
    std::vector<V3> fn(npoly), N(nv, V3{0, 0, 0});
    for (long long f = 0; f < npoly; ++f) fn[f] = cross(P[tri[3 * f + 1]] - P[tri[3 * f]], P[tri[3 * f + 2]] - P[tri[3 * f]]);
    for (int k = 0; k < 3; ++k)
      for (long long f = 0; f < npoly; ++f) {
        V3& n = N[tri[3 * f + k]];
        n.x += fn[f].x; n.y += fn[f].y; n.z += fn[f].z;
      }
    for (size_t i = 0; i < nv; ++i) {
      double l = length(N[i]);
      if (l == 0) l = 1.0;
      N[i] = {N[i].x / l, N[i].y / l, N[i].z / l};
    }
    const size_t base = tri_verts.size();
    tri_verts.resize(base + (size_t)npoly * 9); tri_normals.resize(base + (size_t)npoly * 9);
    for (long long f = 0; f < npoly; ++f)
      for (int k = 0; k < 3; ++k) {
        const V3 p = P[tri[3 * f + k]], n = N[tri[3 * f + k]];
        float* pv = &tri_verts[base + 9 * f + 3 * k]; float* pn = &tri_normals[base + 9 * f + 3 * k];
        pv[0] = (float)p.x; pv[1] = (float)p.y; pv[2] = (float)p.z;
        pn[0] = (float)n.x; pn[1] = (float)n.y; pn[2] = (float)n.z;
      }
    tri_material.insert(tri_material.end(), (size_t)npoly, mat);
    if (nv) grow_bbox(lo, hi);
  }

  void parse_node(const XmlNode* node, const M4& parent, int depth) {
    if (depth > 128) bad("node hierarchy too deep");
    M4 M = identity();
    for (auto& e : node->kids) {
      if (e->tag == "matrix") {
        std::vector<double> v;
        if (!parse_doubles(e->text, &v)) bad("malformed <matrix>");
        M4 t = identity();   // short matrices are padded from the identity (CBgems.dae ships a 15-entry camera matrix)
        for (size_t i = 0; i < v.size() && i < 16; ++i) t.m[i / 4][i % 4] = v[i];
        M = t;
        break;
      }
      if (e->tag == "translate") {
        std::vector<double> v = floats_of(e.get(), "translate", 3);
        M4 T = identity(); T.m[0][3] = v[0]; T.m[1][3] = v[1]; T.m[2][3] = v[2];
        M = mul(T, M);
      }
      // <rotate>/<scale>: mis-read by the reference parser (collada.cpp:259-321: scale reads y twice, never z);
      // none of the bundled path-tracer scenes use them and they are ignored here as in the converter
    }
    M = mul(parent, M);
    for (auto& ch : node->kids) if (ch->tag == "node") parse_node(ch.get(), M, depth + 1);
    const XmlNode* icam = node->child("instance_camera");
    const XmlNode* ilight = node->child("instance_light");
    const XmlNode* igeom = node->child("instance_geometry");
    if (icam) {
      const XmlNode* cam = by_url(icam, "url", "instance_camera");
      const XmlNode* persp = cam->find("optics/technique_common/perspective");
      if (!persp) bad("camera without a perspective block");
      const XmlNode* xf = persp->child("xfov"); const XmlNode* yf = persp->child("yfov");
      hfov = xf ? floats_of(xf, "xfov", 1)[0] : 50.0;
      vfov = yf ? floats_of(yf, "yfov", 1)[0] : 35.0;
      if (!yf) {
        const double PI = 3.14159265358979323846;
        const double ar = floats_of(persp->child("aspect_ratio"), "aspect_ratio", 1)[0];
        vfov = 2 * (std::atan(std::tan((0.5 * hfov) * (PI / 180.0)) / ar) * (180.0 / PI));
      }
      V3 d = xf_point(M, {0, 0, -1});   // point transform, translation included: Application::load quirk
      const double l = length(d);
      cam_dir[0] = d.x / l; cam_dir[1] = d.y / l; cam_dir[2] = d.z / l;
    } else if (ilight) {
      const XmlNode* light = by_url(ilight, "url", "instance_light");
      const XmlNode* t = technique_cmu462(light);
      if (!t) t = technique_common(light);
      if (!t || t->kids.empty()) bad("light without a technique");
      const XmlNode* first = t->kids[0].get();
      std::vector<double> col = floats_of(first->child("color"), "color", 3);
      const V3 pos = xf_point(M, {0, 0, 0});
      V3 dirn = xf_point(M, {0, 0, -1}) - pos;
      const double dl = length(dirn);
      dirn = {dirn.x / dl, dirn.y / dl, dirn.z / dl};
      const V3 up = {0, 1, 0}, ldir = {0, 0, -1};
      const V3 dim_y = xf_point(M, up) - pos;
      const V3 dim_x = xf_point(M, cross(up, ldir)) - pos;
      int kind = -1;
      if (first->tag == "area") kind = B2RT_LIGHT_AREA;
      else if (first->tag == "point") kind = B2RT_LIGHT_POINT;
      else if (first->tag == "directional") kind = B2RT_LIGHT_DIRECTIONAL;
      if (kind >= 0) {   // ambient / hemisphere / spot lights are outside this path: skipped
        b2rt_light L;
        memset(&L, 0, sizeof L);
        L.kind = kind;
        for (int i = 0; i < 3; ++i) L.radiance[i] = (float)col[i];
        L.position[0] = (float)pos.x; L.position[1] = (float)pos.y; L.position[2] = (float)pos.z;
        L.direction[0] = (float)dirn.x; L.direction[1] = (float)dirn.y; L.direction[2] = (float)dirn.z;
        L.dim_x[0] = (float)dim_x.x; L.dim_x[1] = (float)dim_x.y; L.dim_x[2] = (float)dim_x.z;
        L.dim_y[0] = (float)dim_y.x; L.dim_y[1] = (float)dim_y.y; L.dim_y[2] = (float)dim_y.z;
        lights.push_back(L);
      }
    } else if (igeom) {
      const XmlNode* geom = by_url(igeom, "url", "instance_geometry");
      if (const XmlNode* mesh = geom->child("mesh")) {
        add_mesh(mesh, M, material_for(node));
      } else if (geom->child("extra")) {
        const XmlNode* t = technique_cmu462(geom);
        const XmlNode* rad = t ? t->find("sphere/radius") : nullptr;
        if (!rad) bad("geometry <extra> without a CMU462 sphere");
        const double r = floats_of(rad, "radius", 1)[0];
        const V3 c = xf_point(M, {0, 0, 0});
        const double s = length(V3{M.m[0][0], M.m[1][0], M.m[2][0]});
        const double rs = r * s;
        spheres.push_back((float)c.x); spheres.push_back((float)c.y); spheres.push_back((float)c.z); spheres.push_back((float)rs);
        sphere_material.push_back(material_for(node));
        grow_bbox({c.x - rs, c.y - rs, c.z - rs}, {c.x + rs, c.y + rs, c.z + rs});
      }
    }
  }

  void load(const XmlNode* root) {
    // id -> element; a later duplicate id replaces the earlier one, as in the converter's dict
    struct Idx { static void go(const XmlNode* n, std::map<std::string, const XmlNode*>* m) {
      if (const std::string* id = n->attr("id")) (*m)[*id] = n;
      for (auto& k : n->kids) go(k.get(), m);
    } };
    Idx::go(root, &ids);
    const XmlNode* ua = root->find("asset/up_axis");
    if (!ua) bad("no <asset>/<up_axis>");
    std::string up = ua->text;
    up.erase(0, up.find_first_not_of(" \t\r\n"));
    up.erase(up.find_last_not_of(" \t\r\n") + 1);
    M4 G = identity();
    if (up == "X_UP") { G.m[0][0] = 0; G.m[0][1] = 1; G.m[1][0] = 1; G.m[1][1] = 0; G.m[2][2] = -1; }
    else if (up == "Z_UP") { G.m[1][1] = 0; G.m[1][2] = 1; G.m[2][1] = 1; G.m[2][2] = 0; G.m[0][0] = -1; }
    const XmlNode* ivs = root->find("scene/instance_visual_scene");
    if (!ivs) bad("no <scene>/<instance_visual_scene>");
    const XmlNode* vscene = by_url(ivs, "url", "instance_visual_scene");
    for (auto& n : vscene->kids) if (n->tag == "node") parse_node(n.get(), G, 0);
  }
};

}  // namespace

extern "C" int b2rt_load_dae(const char* path, b2rt_scene_file** out) {
  using b2rt::set_error;
  if (!path || !out) { set_error("null argument"); return B2RT_ERR_INVALID; }
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) { set_error(std::string("cannot open ") + path); return B2RT_ERR_IO; }
  std::string buf;
  char chunk[1 << 16];
  size_t n;
  while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) buf.append(chunk, n);
  fclose(f);
  try {
    XmlParser xp(buf.data(), buf.size());
    std::unique_ptr<XmlNode> root = xp.parse_document();
    if (root->tag != "COLLADA") throw LoadError{"root element is <" + root->tag + ">, not <COLLADA>"};
    Loader L;
    L.load(root.get());
    if (L.tri_verts.empty() && L.spheres.empty()) throw LoadError{"scene has no geometry"};
    const float cd[3] = {(float)L.cam_dir[0], (float)L.cam_dir[1], (float)L.cam_dir[2]};
    const float bb[6] = {(float)L.bmin[0], (float)L.bmin[1], (float)L.bmin[2], (float)L.bmax[0], (float)L.bmax[1], (float)L.bmax[2]};
    *out = b2rt::scene_file_from_parts(std::move(L.tri_verts), std::move(L.tri_normals), std::move(L.tri_material), std::move(L.spheres),
                                       std::move(L.sphere_material), std::move(L.materials), std::move(L.lights), cd, (float)L.hfov,
                                       (float)L.vfov, bb);
    return B2RT_OK;
  } catch (const XmlError& e) {
    set_error(std::string(path) + ": " + e.msg);
  } catch (const LoadError& e) {
    set_error(std::string(path) + ": " + e.msg);
  } catch (const std::exception& e) {
    set_error(std::string(path) + ": " + e.what());
  }
  return B2RT_ERR_IO;
}
