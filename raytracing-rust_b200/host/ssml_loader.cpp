// Host-side scene ingest for ptb200: a C++ statement of what the reference's `loader` crate accepts and
// produces, flattened to the POD arrays of include/ptb200.h. No GPU code here.
//
// Reference behaviour followed (crates/loader/src/...):
//   parser.rs:112-197   grammar: [#ver1] { kind [name] ( key value<EOL> ... ) }, values = 3|2|1 doubles | text
//   lib.rs:196-243      load order: textures -> materials -> first camera -> first sky -> primitives -> meshes
//   lib.rs:138-165      Properties accessors (Num1 auto-casts to Vec3/Vec2; float() takes Num1 only)
//   lib.rs:344-399      implicit __DEFAULT_TEX (solid 1.0) and __DEFAULT_MAT (lambertian, albedo 0.25)
//   misc.rs:6-38        camera / sky defaults (aspect fixed 16/9, sampler_res 100x100)
//   textures.rs, materials.rs, primitives.rs, meshes.rs, obj.rs   per-object keys and defaults
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/ptb200.h"

namespace {

thread_local std::string g_host_error;

struct Value {
  enum Kind { NUM1, NUM2, NUM3, TEXT } kind = TEXT;
  float n[3] = {0, 0, 0};
  std::string text;
};
enum ObjKind { K_CAMERA, K_MATERIAL, K_PRIMITIVE, K_SKY, K_TEXTURE, K_MESH, K_OTHER };
struct Object {
  ObjKind kind = K_OTHER;
  bool has_name = false;
  std::string name;
  std::map<std::string, Value> values;  // HashMap semantics: a repeated key keeps the last value
};

// ---- nom-equivalent scanner ------------------------------------------------------------------------------
struct Scanner {
  const std::string& s;
  size_t i = 0;
  explicit Scanner(const std::string& str) : s(str) {}
  bool eof() const { return i >= s.size(); }
  void multispace0() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\r' || s[i] == '\n')) ++i; }
  size_t space0() { size_t b = i; while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) ++i; return i - b; }
  bool tag(const char* t) {
    size_t n = std::strlen(t);
    if (s.compare(i, n, t) == 0) { i += n; return true; }
    return false;
  }
  bool line_ending() {
    if (i < s.size() && s[i] == '\n') { ++i; return true; }
    if (i + 1 < s.size() && s[i] == '\r' && s[i + 1] == '\n') { i += 2; return true; }
    return false;
  }
  // parser.rs:112-117
  bool identifier(std::string& out) {
    size_t b = i;
    if (i < s.size() && (std::isalpha((unsigned char)s[i]) || s[i] == '_')) {
      ++i;
      while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_')) ++i;
      out = s.substr(b, i - b);
      return true;
    }
    return false;
  }
  // nom::number::complete::double: [+-]? (digits [. digits*] | . digits+) ([eE][+-]?digits+)?  |  nan | inf | infinity
  bool number(double& out) {
    size_t b = i, j = i;
    if (j < s.size() && (s[j] == '+' || s[j] == '-')) ++j;
    size_t digits_b = j;
    while (j < s.size() && std::isdigit((unsigned char)s[j])) ++j;
    size_t int_digits = j - digits_b, frac_digits = 0;
    if (j < s.size() && s[j] == '.') {
      size_t k = j + 1;
      while (k < s.size() && std::isdigit((unsigned char)s[k])) ++k;
      frac_digits = k - (j + 1);
      if (int_digits > 0 || frac_digits > 0) j = k;
    }
    if (int_digits == 0 && frac_digits == 0) {
      // exceptions, case-insensitive, no sign handled by nom here either
      auto ci = [&](const char* w) {
        size_t n = std::strlen(w);
        if (b + n > s.size()) return false;
        for (size_t k = 0; k < n; ++k)
          if (std::tolower((unsigned char)s[b + k]) != w[k]) return false;
        return true;
      };
      if (ci("nan")) { i = b + 3; out = std::nan(""); return true; }
      if (ci("infinity")) { i = b + 8; out = INFINITY; return true; }
      if (ci("inf")) { i = b + 3; out = INFINITY; return true; }
      return false;
    }
    if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
      size_t k = j + 1;
      if (k < s.size() && (s[k] == '+' || s[k] == '-')) ++k;
      size_t eb = k;
      while (k < s.size() && std::isdigit((unsigned char)s[k])) ++k;
      if (k > eb) j = k;
    }
    out = std::strtod(s.substr(b, j - b).c_str(), nullptr);
    i = j;
    return true;
  }
};

// parser.rs:119-131 — alt((3 doubles, 2 doubles, 1 double, rest of line))
bool parse_value(Scanner& sc, Value& v) {
  size_t start = sc.i;
  double d[3];
  int got = 0;
  for (; got < 3; ++got) {
    size_t save = sc.i;
    sc.space0();
    if (!sc.number(d[got])) { sc.i = save; break; }
  }
  if (got > 0) {
    v.kind = got == 3 ? Value::NUM3 : (got == 2 ? Value::NUM2 : Value::NUM1);
    for (int k = 0; k < got; ++k) v.n[k] = (float)d[k];
    return true;
  }
  sc.i = start;
  sc.space0();
  size_t b = sc.i;
  while (sc.i < sc.s.size() && sc.s[sc.i] != '\n' && sc.s[sc.i] != '\r') ++sc.i;
  v.kind = Value::TEXT;
  v.text = sc.s.substr(b, sc.i - b);
  return true;
}

// parser.rs:133-168
bool parse_object(Scanner& sc, Object& o) {
  size_t save = sc.i;
  if (sc.tag("camera")) o.kind = K_CAMERA;
  else if (sc.tag("material")) o.kind = K_MATERIAL;
  else if (sc.tag("primitive")) o.kind = K_PRIMITIVE;
  else if (sc.tag("sky")) o.kind = K_SKY;
  else if (sc.tag("texture")) o.kind = K_TEXTURE;
  else if (sc.tag("mesh")) o.kind = K_MESH;
  else return false;
  {  // opt(preceded(space1, identifier))
    size_t s2 = sc.i;
    if (sc.space0() >= 1 && sc.identifier(o.name)) o.has_name = true;
    else sc.i = s2;
  }
  sc.multispace0();
  if (!sc.tag("(")) { sc.i = save; return false; }
  sc.multispace0();
  for (;;) {  // many0(terminated(keyvalue, line_ending))
    size_t kv = sc.i;
    std::string key;
    Value val;
    sc.space0();
    if (!sc.identifier(key) || sc.space0() < 1 || !parse_value(sc, val) || !sc.line_ending()) { sc.i = kv; break; }
    o.values[key] = val;
  }
  sc.multispace0();
  if (!sc.tag(")")) { sc.i = save; return false; }
  sc.multispace0();
  return true;
}

// parser.rs:170-197
bool parse_scene(const std::string& src, std::vector<Object>& out) {
  Scanner sc(src);
  {
    size_t save = sc.i;
    sc.multispace0();
    if (sc.tag("#ver1")) sc.multispace0();
    else sc.i = save;
  }
  for (;;) {
    size_t save = sc.i;
    sc.multispace0();
    Object o;
    if (!parse_object(sc, o)) { sc.i = save; break; }
    out.push_back(o);
  }
  return sc.eof();  // unparsed trailing input => ParseError (parser.rs:188-197)
}

// ---- Properties (lib.rs:103-178) --------------------------------------------------------------------------
struct Props {
  const Object& o;
  explicit Props(const Object& obj) : o(obj) {}
  const Value* get(const char* k) const {
    auto it = o.values.find(k);
    return it == o.values.end() ? nullptr : &it->second;
  }
  bool vec3(const char* k, ptb_vec3& out) const {
    const Value* v = get(k);
    if (!v) return false;
    if (v->kind == Value::NUM3) { out = {v->n[0], v->n[1], v->n[2]}; return true; }
    if (v->kind == Value::NUM1) { out = {v->n[0] * 1.0f, v->n[0] * 1.0f, v->n[0] * 1.0f}; return true; }
    return false;
  }
  bool vec2(const char* k, float out[2]) const {
    const Value* v = get(k);
    if (!v) return false;
    if (v->kind == Value::NUM2) { out[0] = v->n[0]; out[1] = v->n[1]; return true; }
    if (v->kind == Value::NUM1) { out[0] = out[1] = v->n[0]; return true; }
    return false;
  }
  bool flt(const char* k, float& out) const {
    const Value* v = get(k);
    if (v && v->kind == Value::NUM1) { out = v->n[0]; return true; }
    return false;
  }
  const std::string* text(const char* k) const {
    const Value* v = get(k);
    return (v && v->kind == Value::TEXT) ? &v->text : nullptr;
  }
};

inline ptb_vec3 V(float x, float y, float z) { return ptb_vec3{x, y, z}; }
inline ptb_vec3 sub(ptb_vec3 a, ptb_vec3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
inline ptb_vec3 mul(ptb_vec3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
inline ptb_vec3 smul(float s, ptb_vec3 a) { return V(s * a.x, s * a.y, s * a.z); }
inline ptb_vec3 divs(ptb_vec3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
inline float dot(ptb_vec3 a, ptb_vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline ptb_vec3 cross(ptb_vec3 a, ptb_vec3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline ptb_vec3 normalised(ptb_vec3 a) { return divs(a, std::sqrt(dot(a, a))); }

// Rust `f32 as usize` saturating cast, then narrowed to the u32 the ABI carries
inline uint32_t sat_u32(float f) {
  if (!(f > 0.0f)) return 0;
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)f;
}

}  // namespace

struct ptb_host_scene {
  std::vector<ptb_sphere> spheres;
  std::vector<ptb_triangle> triangles;
  std::vector<ptb_material> materials;
  std::vector<ptb_texture> textures;
  struct TexData { uint32_t width = 0, height = 0; std::vector<float> data; };
  std::map<uint32_t, TexData> texture_data;  // ImageTexture pixels / Perlin tables, keyed by texture index
  ptb_camera camera{};
  ptb_sky sky{};
  std::vector<std::string> warnings;
};

namespace {

struct Loader {
  ptb_host_scene* sc;
  std::string base_dir;
  std::map<std::string, uint32_t> tex_lookup, mat_lookup;
  int32_t err = PTB_OK;

  int32_t fail(int32_t code, const std::string& msg) {
    g_host_error = msg;
    err = code;
    return code;
  }

  // textures.rs:5-84
  int32_t load_texture(const Object& o, ptb_texture& t, ptb_host_scene::TexData& td) {
    Props p(o);
    const std::string* kind = p.text("type");
    if (!kind) return fail(PTB_ERR_MISSING, "missing required type for object (texture)");
    t = ptb_texture{};
    if (*kind == "checkered" || *kind == "lerp") {
      t.kind = *kind == "lerp" ? PTB_TEX_LERP : PTB_TEX_CHECKERED;
      if (!p.vec3("primary", t.a)) t.a = V(1, 1, 1);
      if (!p.vec3("secondary", t.b)) t.b = V(0, 0, 0);
    } else if (*kind == "solid") {
      t.kind = PTB_TEX_SOLID;
      if (!p.vec3("colour", t.a)) t.a = smul(0.5f, V(1, 1, 1));
    } else if (*kind == "image") {
      // textures.rs:50-60 + ImageTexture::new (implementations/src/textures/mod.rs:208-245)
      const std::string* fn = p.text("filename");
      if (!fn) return fail(PTB_ERR_MISSING, "missing required value for object: filename");
      t.kind = PTB_TEX_IMAGE;
      std::string path = *fn;
      {
        std::ifstream probe(path, std::ios::binary);
        if (!probe.good() && !base_dir.empty() && !path.empty() && path[0] != '/')
          path = base_dir + "/" + *fn;  // extension: also resolve relative to the scene file
      }
      float* px = nullptr;
      int32_t rc = ptb_image_load(path.c_str(), &td.width, &td.height, &px);
      if (rc != PTB_OK) return fail(rc, std::string("image texture: ") + ptb_image_last_error());
      td.data.assign(px, px + (size_t)td.width * td.height * 3);
      ptb_image_free(px);
    } else if (*kind == "perlin") {
      // textures.rs:62-67 + Perlin::new (textures/mod.rs:90-112); the reference seeds the tables from OS entropy, here
      // they are a pure function of the (optional, extension) `seed` key and the texture's position in the file
      t.kind = PTB_TEX_PERLIN;
      float seed = 0.0f;
      p.flt("seed", seed);
      td.data.resize(PTB_PERLIN_TABLE_WORDS);
      ptb_perlin_tables(0x9E3779B97F4A7C15ull ^ ((uint64_t)(int64_t)seed << 20) ^ (uint64_t)sc->textures.size(), td.data.data());
    } else {
      return fail(PTB_ERR_MISSING, "required a known value for texture type, found '" + *kind + "'");
    }
    return PTB_OK;
  }

  uint32_t texture_for(const Props& p) {  // props.texture("texture").unwrap_or_else(default_texture)
    const std::string* name = p.text("texture");
    if (name) {
      auto it = tex_lookup.find(*name);
      if (it != tex_lookup.end()) return it->second;
    }
    return tex_lookup["__DEFAULT_TEX"];
  }
  uint32_t material_named(const std::string& name, bool& found) {
    auto it = mat_lookup.find(name);
    found = it != mat_lookup.end();
    return found ? it->second : mat_lookup["__DEFAULT_MAT"];
  }
  uint32_t material_for(const Props& p) {  // props.scatter("material").unwrap_or_else(default_scatter)
    const std::string* name = p.text("material");
    bool found;
    if (name) return material_named(*name, found);
    return mat_lookup["__DEFAULT_MAT"];
  }

  // materials.rs:6-111
  int32_t load_material(const Object& o, ptb_material& m) {
    Props p(o);
    const std::string* kind = p.text("type");
    if (!kind) return fail(PTB_ERR_MISSING, "missing required type for object (material)");
    m = ptb_material{};
    m.texture = texture_for(p);
    m.ior = V(1, 1, 1);
    if (*kind == "emissive") {
      m.kind = PTB_MAT_EMIT;
      if (!p.flt("strength", m.param)) m.param = 1.5f;
    } else if (*kind == "lambertian") {
      m.kind = PTB_MAT_LAMBERTIAN;
      if (!p.flt("albedo", m.param)) m.param = 0.5f;
    } else if (*kind == "reflect") {
      m.kind = PTB_MAT_REFLECT;
      if (!p.flt("fuzz", m.param)) m.param = 0.1f;
    } else if (*kind == "refract") {
      m.kind = PTB_MAT_REFRACT;
      if (!p.flt("eta", m.param)) m.param = 1.5f;
    } else if (*kind == "trowbridge_reitz") {
      m.kind = PTB_MAT_TROWBRIDGE_REITZ;
      float alpha;
      if (!p.flt("alpha", alpha)) alpha = 0.5f;
      m.param = alpha * alpha;  // TrowbridgeReitz::new stores roughness^2 (trowbridge_reitz.rs:17-24)
      if (!p.vec3("ior", m.ior)) m.ior = V(1, 1, 1);
      if (!p.flt("metallic", m.metallic)) m.metallic = 0.0f;
    } else {
      return fail(PTB_ERR_MISSING, "required a known value for material type, found '" + *kind + "'");
    }
    return PTB_OK;
  }

  // meshes.rs:26-103
  int32_t load_cuboid(const Object& o) {
    Props p(o);
    uint32_t mat = material_for(p);
    ptb_vec3 p1, p2;
    if (!p.vec3("point_one", p1)) return fail(PTB_ERR_MISSING, "expected point_one on aacubiod, found nothing");
    if (!p.vec3("point_two", p2)) return fail(PTB_ERR_MISSING, "expected point_two on aacubiod, found nothing");
    ptb_vec3 mn = V(std::fmin(p1.x, p2.x), std::fmin(p1.y, p2.y), std::fmin(p1.z, p2.z));
    ptb_vec3 mx = V(std::fmax(p1.x, p2.x), std::fmax(p1.y, p2.y), std::fmax(p1.z, p2.z));
    const ptb_vec3 pts[8] = {mn, V(mx.x, mn.y, mn.z), V(mx.x, mx.y, mn.z), V(mn.x, mx.y, mn.z),
                             V(mn.x, mn.y, mx.z), V(mx.x, mn.y, mx.z), mx, V(mn.x, mx.y, mx.z)};
    const ptb_vec3 nrm[6] = {V(1, 0, 0), V(-1, 0, 0), V(0, 1, 0), V(0, -1, 0), V(0, 0, 1), V(0, 0, -1)};
    static const int tri[12][4] = {{0, 1, 2, 5}, {0, 2, 3, 5}, {0, 1, 5, 3}, {0, 5, 4, 3}, {1, 2, 5, 0}, {2, 5, 6, 0},
                                   {2, 3, 7, 2}, {2, 6, 7, 2}, {0, 3, 4, 1}, {3, 4, 7, 1}, {4, 5, 6, 4}, {4, 6, 7, 4}};
    for (int t = 0; t < 12; ++t) {
      ptb_triangle T{};
      for (int k = 0; k < 3; ++k) { T.p[k] = pts[tri[t][k]]; T.n[k] = nrm[tri[t][3]]; }
      T.material = mat;
      sc->triangles.push_back(T);
    }
    return PTB_OK;
  }

  // obj.rs:11-65 over the `v / vn / vt / f / usemtl / o` subset of wavefront_obj 10 (triangles; polygons as fans)
  int32_t load_obj(const std::string& path_in, const Object&) {
    std::ifstream f(path_in);
    std::string path = path_in;
    if (!f.good() && !base_dir.empty() && !path_in.empty() && path_in[0] != '/') {
      path = base_dir + "/" + path_in;  // extension: also resolve relative to the scene file
      f.close();
      f.clear();
      f.open(path);
    }
    if (!f.good()) return fail(PTB_ERR_IO, "failed to read obj file '" + path_in + "'");
    std::vector<ptb_vec3> verts, norms;
    bool found;
    uint32_t cur_mat = material_named("default", found);
    std::string line;
    size_t lineno = 0;
    while (std::getline(f, line)) {
      ++lineno;
      if (!line.empty() && line.back() == '\r') line.pop_back();
      std::istringstream ls(line);
      std::string tok;
      if (!(ls >> tok) || tok[0] == '#') continue;
      if (tok == "v" || tok == "vn") {
        double x, y, z;
        if (!(ls >> x >> y >> z)) return fail(PTB_ERR_PARSE, "obj: bad vertex at line " + std::to_string(lineno));
        (tok == "v" ? verts : norms).push_back(V((float)x, (float)y, (float)z));
      } else if (tok == "usemtl") {
        std::string name;
        ls >> name;
        cur_mat = material_named(name, found);
      } else if (tok == "f") {
        std::vector<std::pair<long, long>> idx;  // (vertex, normal), 0-based; normal -1 when absent
        std::string vert;
        while (ls >> vert) {
          long vi = 0, ni = 0;
          bool has_n = false;
          size_t s1 = vert.find('/');
          vi = std::strtol(vert.substr(0, s1).c_str(), nullptr, 10);
          if (s1 != std::string::npos) {
            size_t s2 = vert.find('/', s1 + 1);
            if (s2 != std::string::npos && s2 + 1 < vert.size()) {
              ni = std::strtol(vert.substr(s2 + 1).c_str(), nullptr, 10);
              has_n = true;
            }
          }
          if (vi < 0) vi = (long)verts.size() + vi; else vi -= 1;
          if (has_n) { if (ni < 0) ni = (long)norms.size() + ni; else ni -= 1; }
          else ni = -1;
          idx.push_back(std::make_pair(vi, ni));
        }
        if (idx.size() < 3) continue;
        for (size_t k = 1; k + 1 < idx.size(); ++k) {
          const std::pair<long, long> c[3] = {idx[0], idx[k], idx[k + 1]};
          ptb_triangle T{};
          for (int q = 0; q < 3; ++q) {
            if (c[q].second < 0) return fail(PTB_ERR_PARSE, "Please export obj file with vertex normals!");
            if (c[q].first < 0 || (size_t)c[q].first >= verts.size() || (size_t)c[q].second >= norms.size())
              return fail(PTB_ERR_PARSE, "obj: index out of range at line " + std::to_string(lineno));
            T.p[q] = verts[c[q].first];
            T.n[q] = norms[c[q].second];
          }
          T.material = cur_mat;
          sc->triangles.push_back(T);
        }
      }
      // o / g / s / vt / mtllib: no effect on the flattened triangle list
    }
    return PTB_OK;
  }

  int32_t run(const std::string& text) {
    std::vector<Object> objects;
    if (!parse_scene(text, objects)) return fail(PTB_ERR_PARSE, "failed to parse the scene config");

    // textures (lib.rs:344-370): scene textures in file order, then __DEFAULT_TEX
    for (const Object& o : objects) {
      if (o.kind != K_TEXTURE) continue;
      ptb_texture t;
      ptb_host_scene::TexData td;
      if (load_texture(o, t, td) != PTB_OK) return err;
      sc->textures.push_back(t);
      if (!td.data.empty()) sc->texture_data[(uint32_t)sc->textures.size() - 1] = std::move(td);
      if (o.has_name) {
        if (tex_lookup.count(o.name)) sc->warnings.push_back("Overwrote previous object of name: '" + o.name + "'");
        tex_lookup[o.name] = (uint32_t)sc->textures.size() - 1;
      }
    }
    {
      ptb_texture t{};
      t.kind = PTB_TEX_SOLID;
      t.a = V(1, 1, 1);
      sc->textures.push_back(t);
      tex_lookup["__DEFAULT_TEX"] = (uint32_t)sc->textures.size() - 1;
    }
    // materials (lib.rs:372-399)
    for (const Object& o : objects) {
      if (o.kind != K_MATERIAL) continue;
      ptb_material m;
      if (load_material(o, m) != PTB_OK) return err;
      sc->materials.push_back(m);
      if (o.has_name) {
        if (mat_lookup.count(o.name)) sc->warnings.push_back("Overwrote previous object of name: '" + o.name + "'");
        mat_lookup[o.name] = (uint32_t)sc->materials.size() - 1;
      }
    }
    {
      ptb_material m{};
      m.kind = PTB_MAT_LAMBERTIAN;
      m.texture = tex_lookup["__DEFAULT_TEX"];
      m.param = 0.25f;
      m.ior = V(1, 1, 1);
      sc->materials.push_back(m);
      mat_lookup["__DEFAULT_MAT"] = (uint32_t)sc->materials.size() - 1;
    }
    // camera (lib.rs:280-296, misc.rs:6-18)
    {
      const Object* cam = nullptr;
      for (const Object& o : objects)
        if (o.kind == K_CAMERA) { cam = &o; break; }
      if (!cam) return fail(PTB_ERR_MISSING, "missing required camera object");
      Props p(*cam);
      ptb_vec3 origin, lookat, vup;
      float fov, aperture, focus;
      if (!p.vec3("origin", origin)) origin = V(3, 0, 0);
      if (!p.vec3("lookat", lookat)) lookat = V(0, 0, 0);
      if (!p.vec3("vup", vup)) vup = V(0, 1, 0);
      if (!p.flt("fov", fov)) fov = 40.0f;
      if (!p.flt("aperture", aperture)) aperture = 0.0f;
      if (!p.flt("focus_dis", focus)) focus = 10.0f;
      ptb_camera_make(origin, lookat, vup, fov, 16.0f / 9.0f, aperture, focus, &sc->camera);
    }
    // sky (lib.rs:298-316, misc.rs:20-38)
    {
      const Object* sky = nullptr;
      for (const Object& o : objects)
        if (o.kind == K_SKY) { sky = &o; break; }
      Object empty;
      if (!sky) {
        sc->warnings.push_back("no sky object was provided in scene file, using default");
        sky = &empty;
      }
      Props p(*sky);
      float res[2];
      if (!p.vec2("sampler_res", res)) { res[0] = 100.0f; res[1] = 100.0f; }
      sc->sky.texture = texture_for(p);
      sc->sky.sampler_res_x = sat_u32(res[0]);
      sc->sky.sampler_res_y = sat_u32(res[1]);
    }
    // primitives (lib.rs:401-412, primitives.rs:8-50)
    for (const Object& o : objects) {
      if (o.kind != K_PRIMITIVE) continue;
      Props p(o);
      const std::string* kind = p.text("type");
      if (!kind) return fail(PTB_ERR_MISSING, "missing required type for object (primitive)");
      if (*kind == "sphere") {
        ptb_sphere s{};
        s.material = material_for(p);
        if (!p.flt("radius", s.radius)) s.radius = 1.0f;
        if (!p.vec3("centre", s.center)) return fail(PTB_ERR_MISSING, "expected centre on sphere, found nothing");
        sc->spheres.push_back(s);
      } else if (*kind == "triangle") {
        return fail(PTB_ERR_UNSUPPORTED, "primitive type 'triangle' is todo!() in the reference loader (primitives.rs:42)");
      } else {
        return fail(PTB_ERR_MISSING, "required a known value for primitive type, found '" + *kind + "'");
      }
    }
    // meshes (lib.rs:414-428, meshes.rs:9-24,105-119)
    for (const Object& o : objects) {
      if (o.kind != K_MESH) continue;
      Props p(o);
      const std::string* kind = p.text("type");
      if (!kind) return fail(PTB_ERR_MISSING, "missing required type for object (mesh)");
      if (*kind == "mesh") {
        const std::string* path = p.text("obj");
        if (!path) return fail(PTB_ERR_MISSING, "expected obj on mesh, found nothing");
        if (load_obj(*path, o) != PTB_OK) return err;
      } else if (*kind == "aacuboid") {
        if (load_cuboid(o) != PTB_OK) return err;
      } else {
        return fail(PTB_ERR_MISSING, "required a known value for mesh type, found '" + *kind + "'");
      }
    }
    return PTB_OK;
  }
};

}  // namespace

extern "C" {

const char* ptb_host_last_error(void) { return g_host_error.c_str(); }

// camera.rs:20-53
int32_t ptb_camera_make(ptb_vec3 origin, ptb_vec3 lookat, ptb_vec3 vup, float hfov_deg, float aspect, float /*aperture*/,
                        float focus_dist, ptb_camera* out) {
  if (!out) return PTB_ERR_INVALID;
  const float pi = 3.14159265358979323846f;
  float fov_rad = hfov_deg * (pi / 180.0f);
  float viewport_width = 2.0f * std::tan(fov_rad / 2.0f);
  float viewport_height = viewport_width / aspect;
  ptb_vec3 w = normalised(sub(origin, lookat));
  ptb_vec3 u = normalised(cross(w, vup));
  ptb_vec3 v = cross(u, w);
  ptb_vec3 horizontal = mul(smul(focus_dist, u), viewport_width);
  ptb_vec3 vertical = mul(smul(focus_dist, v), viewport_height);
  out->origin = origin;
  out->horizontal = horizontal;
  out->vertical = vertical;
  out->lower_left = sub(sub(sub(origin, divs(horizontal, 2.0f)), divs(vertical, 2.0f)), smul(focus_dist, w));
  return PTB_OK;
}

int32_t ptb_ssml_load_str(const char* text, const char* base_dir, ptb_host_scene** out) {
  if (!text || !out) return PTB_ERR_INVALID;
  ptb_host_scene* sc = new (std::nothrow) ptb_host_scene();
  if (!sc) return PTB_ERR_OOM;
  Loader L;
  L.sc = sc;
  L.base_dir = base_dir ? base_dir : "";
  int32_t rc;
  try {
    rc = L.run(text);
  } catch (const std::bad_alloc&) {
    g_host_error = "out of host memory while loading scene";
    rc = PTB_ERR_OOM;
  } catch (const std::exception& e) {
    g_host_error = e.what();
    rc = PTB_ERR_INVALID;
  }
  if (rc != PTB_OK) { delete sc; *out = nullptr; return rc; }
  *out = sc;
  return PTB_OK;
}

int32_t ptb_ssml_load_file(const char* path, ptb_host_scene** out) {
  if (!path || !out) return PTB_ERR_INVALID;
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) {
    g_host_error = std::string("failed to load a file for the given reason: ") + path;
    return PTB_ERR_IO;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  std::string p(path);
  size_t slash = p.find_last_of('/');
  std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
  return ptb_ssml_load_str(ss.str().c_str(), dir.c_str(), out);
}

void ptb_host_scene_free(ptb_host_scene* s) { delete s; }
size_t ptb_host_scene_spheres(const ptb_host_scene* s, const ptb_sphere** out) {
  if (out) *out = s->spheres.data();
  return s->spheres.size();
}
size_t ptb_host_scene_triangles(const ptb_host_scene* s, const ptb_triangle** out) {
  if (out) *out = s->triangles.data();
  return s->triangles.size();
}
size_t ptb_host_scene_materials(const ptb_host_scene* s, const ptb_material** out) {
  if (out) *out = s->materials.data();
  return s->materials.size();
}
size_t ptb_host_scene_textures(const ptb_host_scene* s, const ptb_texture** out) {
  if (out) *out = s->textures.data();
  return s->textures.size();
}
size_t ptb_host_scene_texture_data(const ptb_host_scene* s, uint32_t texture, uint32_t* width, uint32_t* height,
                                   const float** data) {
  auto it = s->texture_data.find(texture);
  if (it == s->texture_data.end()) return 0;
  if (width) *width = it->second.width;
  if (height) *height = it->second.height;
  if (data) *data = it->second.data.data();
  return it->second.data.size();
}
int32_t ptb_host_scene_camera(const ptb_host_scene* s, ptb_camera* out) {
  if (!s || !out) return PTB_ERR_INVALID;
  *out = s->camera;
  return PTB_OK;
}
int32_t ptb_host_scene_sky(const ptb_host_scene* s, ptb_sky* out) {
  if (!s || !out) return PTB_ERR_INVALID;
  *out = s->sky;
  return PTB_OK;
}

}  // extern "C"
