// scene.cpp — JSON decoding/encoding and flattening of scene::Scene.  See scene.h.
#include "scene.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <atomic>
#include <map>
#include <sstream>
#include <stdexcept>

namespace scene {
namespace {

// ------------------------------------------------------------------ minimal JSON value + parser
struct JValue {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;   // insertion order; later duplicates win on lookup
};

struct Parser {
    const char* p;
    const char* end;
    [[noreturn]] void err(const char* what) { throw std::runtime_error(std::string("decode scene: ") + what); }
    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    JValue value() {
        ws();
        if (p >= end) err("unexpected end of JSON input");
        JValue v;
        char c = *p;
        if (c == '{') {
            ++p; v.kind = JValue::Object; ws();
            if (p < end && *p == '}') { ++p; return v; }
            for (;;) {
                ws();
                if (p >= end || *p != '"') err("invalid character looking for beginning of object key string");
                std::string k = string();
                ws();
                if (p >= end || *p != ':') err("invalid character after object key");
                ++p;
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                err("invalid character after object key:value pair");
            }
        } else if (c == '[') {
            ++p; v.kind = JValue::Array; ws();
            if (p < end && *p == ']') { ++p; return v; }
            for (;;) {
                v.arr.push_back(value());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                err("invalid character after array element");
            }
        } else if (c == '"') {
            v.kind = JValue::String; v.str = string();
        } else if (c == 't' && end - p >= 4 && !std::strncmp(p, "true", 4)) { p += 4; v.kind = JValue::Bool; v.b = true; }
        else if (c == 'f' && end - p >= 5 && !std::strncmp(p, "false", 5)) { p += 5; v.kind = JValue::Bool; v.b = false; }
        else if (c == 'n' && end - p >= 4 && !std::strncmp(p, "null", 4)) { p += 4; v.kind = JValue::Null; }
        else if (c == '-' || (c >= '0' && c <= '9')) {
            const char* s = p;
            if (*p == '-') ++p;
            if (p >= end || !std::isdigit((unsigned char)*p)) err("invalid number literal");
            while (p < end && std::isdigit((unsigned char)*p)) ++p;
            if (p < end && *p == '.') { ++p; if (p >= end || !std::isdigit((unsigned char)*p)) err("invalid number literal"); while (p < end && std::isdigit((unsigned char)*p)) ++p; }
            if (p < end && (*p == 'e' || *p == 'E')) { ++p; if (p < end && (*p == '+' || *p == '-')) ++p; if (p >= end || !std::isdigit((unsigned char)*p)) err("invalid number literal"); while (p < end && std::isdigit((unsigned char)*p)) ++p; }
            v.kind = JValue::Number;
            v.num = std::strtod(std::string(s, p).c_str(), nullptr);   // correctly rounded, like strconv.ParseFloat
        } else err("invalid character looking for beginning of value");
        return v;
    }
    static void utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    unsigned hex4() {
        if (end - p < 4) err("invalid unicode escape");
        unsigned v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else err("invalid unicode escape");
        }
        return v;
    }
    std::string string() {
        ++p;   // opening quote
        std::string o;
        while (p < end && *p != '"') {
            char c = *p++;
            if (c == '\\') {
                if (p >= end) err("unexpected end of JSON input");
                char e = *p++;
                switch (e) {
                case '"': o += '"'; break; case '\\': o += '\\'; break; case '/': o += '/'; break;
                case 'b': o += '\b'; break; case 'f': o += '\f'; break; case 'n': o += '\n'; break;
                case 'r': o += '\r'; break; case 't': o += '\t'; break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2;
                        unsigned lo = hex4();
                        if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        else { utf8(o, 0xFFFD); cp = lo; }
                    }
                    utf8(o, cp);
                    break;
                }
                default: err("invalid escape in string");
                }
            } else if ((unsigned char)c < 0x20) err("invalid control character in string literal");
            else o += c;
        }
        if (p >= end) err("unexpected end of JSON input");
        ++p;
        return o;
    }
};

bool ieq(const std::string& a, const char* b) {
    size_t n = std::strlen(b);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; i++) if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}

// encoding/json field matching: every key of the object is applied in order to the field whose name matches
// exactly, else case-insensitively; later keys overwrite earlier ones; unknown keys are ignored (io.go:18 does
// not call DisallowUnknownFields).
struct Fields {
    const JValue& o;
    template <class F> void each(const char* name, F&& f) const {
        for (auto& kv : o.obj)
            if (kv.first == name || ieq(kv.first, name)) f(kv.second);
    }
};
[[noreturn]] void type_err(const char* field, const char* want) {
    throw std::runtime_error(std::string("decode scene: json: cannot unmarshal value into field ") + field + " of type " + want);
}
void get(const JValue& o, const char* k, double& out) {
    Fields{o}.each(k, [&](const JValue& v) { if (v.kind == JValue::Number) out = v.num; else if (v.kind != JValue::Null) type_err(k, "float64"); });
}
void get(const JValue& o, const char* k, int& out) {
    Fields{o}.each(k, [&](const JValue& v) {
        if (v.kind == JValue::Number) { if (v.num != std::floor(v.num)) type_err(k, "int"); out = (int)v.num; }
        else if (v.kind != JValue::Null) type_err(k, "int");
    });
}
void get(const JValue& o, const char* k, bool& out) {
    Fields{o}.each(k, [&](const JValue& v) { if (v.kind == JValue::Bool) out = v.b; else if (v.kind != JValue::Null) type_err(k, "bool"); });
}
void get(const JValue& o, const char* k, std::string& out) {
    Fields{o}.each(k, [&](const JValue& v) { if (v.kind == JValue::String) out = v.str; else if (v.kind != JValue::Null) type_err(k, "string"); });
}
void get(const JValue& o, const char* k, Vec3& out) {
    Fields{o}.each(k, [&](const JValue& v) {
        if (v.kind == JValue::Object) { get(v, "x", out.X); get(v, "y", out.Y); get(v, "z", out.Z); }
        else if (v.kind != JValue::Null) type_err(k, "scene.Vec3");
    });
}
void get(const JValue& o, const char* k, Color& out) {
    Fields{o}.each(k, [&](const JValue& v) {
        if (v.kind == JValue::Object) { get(v, "r", out.R); get(v, "g", out.G); get(v, "b", out.B); }
        else if (v.kind != JValue::Null) type_err(k, "scene.Color");
    });
}

std::unique_ptr<Scene> decode(const JValue& root) {
    if (root.kind != JValue::Object) throw std::runtime_error("decode scene: json: cannot unmarshal value into Go value of type scene.Scene");
    auto sc = std::make_unique<Scene>();
    get(root, "name", sc->Name);
    Fields{root}.each("camera", [&](const JValue& v) {
        if (v.kind == JValue::Null) return;
        if (v.kind != JValue::Object) type_err("camera", "scene.Camera");
        Camera& c = sc->Cam;
        get(v, "position", c.Position); get(v, "target", c.Target); get(v, "up", c.Up);
        get(v, "fov", c.FOV); get(v, "aperture", c.Aperture); get(v, "focus_dist", c.FocusDist); get(v, "aspect_ratio", c.AspectRatio);
    });
    Fields{root}.each("objects", [&](const JValue& v) {
        if (v.kind == JValue::Null) { sc->Objects.clear(); return; }
        if (v.kind != JValue::Array) type_err("objects", "[]scene.Object");
        sc->Objects.clear();
        for (auto& e : v.arr) {
            Object o;
            if (e.kind == JValue::Object) {
                get(e, "id", o.ID); get(e, "type", o.Type); get(e, "position", o.Position); get(e, "size", o.Size);
                get(e, "material_id", o.MaterialID);
                Fields{e}.each("mesh", [&](const JValue& mv) {          // EXTENSION
                    if (mv.kind != JValue::Object) return;
                    auto mesh = std::make_shared<MeshData>();
                    Fields{mv}.each("vertices", [&](const JValue& a) {
                        if (a.kind != JValue::Array) type_err("mesh.vertices", "[]float32");
                        mesh->vertices.clear();
                        for (auto& x : a.arr) { if (x.kind != JValue::Number) type_err("mesh.vertices", "[]float32"); mesh->vertices.push_back((float)x.num); }
                    });
                    Fields{mv}.each("triangles", [&](const JValue& a) {
                        if (a.kind != JValue::Array) type_err("mesh.triangles", "[]uint32");
                        mesh->triangles.clear();
                        for (auto& x : a.arr) { if (x.kind != JValue::Number || x.num < 0) type_err("mesh.triangles", "[]uint32"); mesh->triangles.push_back((uint32_t)x.num); }
                    });
                    Fields{mv}.each("heightfield", [&](const JValue& h) {
                        if (h.kind != JValue::Object) return;
                        int seed = 0;
                        get(h, "nx", mesh->nx); get(h, "nz", mesh->nz); get(h, "seed", seed); get(h, "octaves", mesh->octaves);
                        get(h, "amplitude", mesh->amplitude); get(h, "frequency", mesh->frequency);
                        mesh->seed = (uint32_t)seed;
                        mesh->generated = true;
                        GenerateHeightfield(*mesh);
                    });
                    if (mesh->vertices.size() % 3 || mesh->triangles.size() % 3) type_err("mesh", "triples");
                    for (uint32_t idx : mesh->triangles) if ((size_t)idx >= mesh->vertices.size() / 3) type_err("mesh.triangles", "vertex index in range");
                    o.Mesh = mesh;
                });
            } else if (e.kind != JValue::Null) type_err("objects", "scene.Object");
            sc->Objects.push_back(std::move(o));
        }
    });
    Fields{root}.each("materials", [&](const JValue& v) {
        if (v.kind == JValue::Null) { sc->Materials.clear(); return; }
        if (v.kind != JValue::Array) type_err("materials", "[]scene.Material");
        sc->Materials.clear();
        for (auto& e : v.arr) {
            Material m;
            if (e.kind == JValue::Object) {
                get(e, "id", m.ID); get(e, "type", m.Type); get(e, "albedo", m.Albedo); get(e, "rough", m.Rough);
                get(e, "ior", m.IOR); get(e, "emit", m.Emit); get(e, "power", m.Power); get(e, "absorption", m.Absorption);
                get(e, "smoothness", m.Smoothness); get(e, "reflectivity", m.Reflectivity); get(e, "tint", m.Tint);
                get(e, "absorption_scale", m.AbsorptionScale);
            } else if (e.kind != JValue::Null) type_err("materials", "scene.Material");
            sc->Materials.push_back(std::move(m));
        }
    });
    Fields{root}.each("settings", [&](const JValue& v) {
        if (v.kind == JValue::Null) return;
        if (v.kind != JValue::Object) type_err("settings", "scene.RenderSettings");
        get(v, "width", sc->Settings.Width); get(v, "height", sc->Settings.Height);
        get(v, "samples_per_px", sc->Settings.SamplesPerPx); get(v, "max_depth", sc->Settings.MaxDepth);
    });
    get(root, "background", sc->Background);
    Fields{root}.each("sky", [&](const JValue& v) {
        if (v.kind == JValue::Null) { sc->SkyPtr.reset(); return; }
        if (v.kind != JValue::Object) type_err("sky", "scene.Sky");
        if (!sc->SkyPtr) sc->SkyPtr = std::make_unique<Sky>();
        get(v, "type", sc->SkyPtr->Type); get(v, "color", sc->SkyPtr->SkyColor);
        get(v, "horizon", sc->SkyPtr->Horizon); get(v, "zenith", sc->SkyPtr->Zenith);
    });
    Fields{root}.each("fog", [&](const JValue& v) {
        if (v.kind == JValue::Null) { sc->FogPtr.reset(); return; }
        if (v.kind != JValue::Object) type_err("fog", "scene.Fog");
        if (!sc->FogPtr) sc->FogPtr = std::make_unique<Fog>();
        Fog& f = *sc->FogPtr;
        get(v, "density", f.Density); get(v, "color", f.FogColor); get(v, "scatter", f.Scatter); get(v, "sigma_s", f.SigmaS);
        get(v, "sigma_a", f.SigmaA); get(v, "g", f.G); get(v, "hetero_strength", f.HeteroStrength);
        get(v, "noise_scale", f.NoiseScale); get(v, "noise_octaves", f.NoiseOctaves); get(v, "affect_sky", f.AffectSky);
        get(v, "gpu_volumetric", f.GPUVolumetric);
    });
    return sc;
}

// ------------------------------------------------------------------ encoder (Go json.Encoder with SetIndent("", "  "))
std::string fmt_float(double v) {   // shortest representation that round-trips, Go 'g'-like exponent rule
    if (v == 0) return std::signbit(v) ? "-0" : "0";
    char buf[40];
    for (int prec = 1; prec <= 17; prec++) {
        std::snprintf(buf, sizeof buf, "%.*e", prec - 1, v);
        if (std::strtod(buf, nullptr) == v) break;
    }
    // buf is d.ddddde[+-]XX ; Go uses exponent form iff exp < -6 || exp >= 21
    std::string s(buf);
    size_t epos = s.find('e');
    int exp = std::atoi(s.c_str() + epos + 1);
    std::string mant = s.substr(0, epos);
    bool neg = mant[0] == '-';
    if (neg) mant = mant.substr(1);
    std::string digits;
    for (char c : mant) if (c != '.') digits += c;
    std::string out;
    if (exp < -6 || exp >= 21) {
        out = digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        char eb[16];
        std::snprintf(eb, sizeof eb, "e%c%02d", exp < 0 ? '-' : '+', std::abs(exp));
        out += eb;
    } else if (exp < 0) {
        out = "0." + std::string(-exp - 1, '0') + digits;
    } else if ((int)digits.size() <= exp + 1) {
        out = digits + std::string(exp + 1 - digits.size(), '0');
    } else {
        out = digits.substr(0, exp + 1) + "." + digits.substr(exp + 1);
    }
    return neg ? "-" + out : out;
}
std::string fmt_string(const std::string& s) {   // Go escapes <, >, & and control characters (EscapeHTML default on)
    std::string o = "\"";
    for (unsigned char c : s) {
        switch (c) {
        case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break;
        case '\r': o += "\\r"; break; case '\t': o += "\\t"; break;
        case '<': o += "\\u003c"; break; case '>': o += "\\u003e"; break; case '&': o += "\\u0026"; break;
        default:
            if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; }
            else o += (char)c;
        }
    }
    return o + "\"";
}
struct Enc {
    std::string out;
    int depth = 0;
    std::vector<bool> first;
    void nl() { out += "\n"; out.append((size_t)depth * 2, ' '); }
    void open(char c) { out += c; depth++; first.push_back(true); }
    void close(char c) { depth--; bool was_empty = first.back(); first.pop_back(); if (!was_empty) nl(); out += c; }
    void key(const char* k) { if (!first.back()) out += ","; first.back() = false; nl(); out += fmt_string(k); out += ": "; }
    void elem() { if (!first.back()) out += ","; first.back() = false; nl(); }
    void num(const char* k, double v) { key(k); out += fmt_float(v); }
    void integer(const char* k, int v) { key(k); out += std::to_string(v); }
    void boolean(const char* k, bool v) { key(k); out += v ? "true" : "false"; }
    void str(const char* k, const std::string& v) { key(k); out += fmt_string(v); }
    void vec3(const char* k, const Vec3& v) { key(k); open('{'); num("x", v.X); num("y", v.Y); num("z", v.Z); close('}'); }
    void color(const char* k, const Color& v) { key(k); open('{'); num("r", v.R); num("g", v.G); num("b", v.B); close('}'); }
};

}  // namespace

std::unique_ptr<Scene> Parse(const std::string& text) {
    Parser ps{text.data(), text.data() + text.size()};
    JValue root = ps.value();   // json.Decoder.Decode reads ONE value; trailing data is not an error (io.go:18)
    return decode(root);
}

std::unique_ptr<Scene> Load(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("open scene: open " + path + ": " + std::strerror(errno));
    std::stringstream ss;
    ss << f.rdbuf();
    return Parse(ss.str());
}

std::string Marshal(const Scene& sc) {
    Enc e;
    e.open('{');
    e.str("name", sc.Name);
    e.key("camera"); e.open('{');
    e.vec3("position", sc.Cam.Position); e.vec3("target", sc.Cam.Target); e.vec3("up", sc.Cam.Up);
    e.num("fov", sc.Cam.FOV); e.num("aperture", sc.Cam.Aperture); e.num("focus_dist", sc.Cam.FocusDist);
    e.num("aspect_ratio", sc.Cam.AspectRatio);
    e.close('}');
    e.key("objects");
    e.open('[');
    for (auto& o : sc.Objects) {
        e.elem(); e.open('{');
        e.str("id", o.ID); e.str("type", o.Type); e.vec3("position", o.Position); e.vec3("size", o.Size);
        e.str("material_id", o.MaterialID);
        if (o.Mesh) {                                                   // EXTENSION (omitted when absent)
            e.key("mesh"); e.open('{');
            if (o.Mesh->generated) {
                e.key("heightfield"); e.open('{');
                e.integer("nx", o.Mesh->nx); e.integer("nz", o.Mesh->nz); e.integer("seed", (int)o.Mesh->seed);
                e.integer("octaves", o.Mesh->octaves); e.num("amplitude", o.Mesh->amplitude); e.num("frequency", o.Mesh->frequency);
                e.close('}');
            } else {
                e.key("vertices"); e.out += "[";
                for (size_t i = 0; i < o.Mesh->vertices.size(); i++) { if (i) e.out += ","; e.out += fmt_float((double)o.Mesh->vertices[i]); }
                e.out += "]";
                e.key("triangles"); e.out += "[";
                for (size_t i = 0; i < o.Mesh->triangles.size(); i++) { if (i) e.out += ","; e.out += std::to_string(o.Mesh->triangles[i]); }
                e.out += "]";
            }
            e.close('}');
        }
        e.close('}');
    }
    e.close(']');
    e.key("materials");
    e.open('[');
    for (auto& m : sc.Materials) {
        e.elem(); e.open('{');
        e.str("id", m.ID); e.str("type", m.Type); e.color("albedo", m.Albedo); e.num("rough", m.Rough); e.num("ior", m.IOR);
        e.color("emit", m.Emit); e.num("power", m.Power); e.color("absorption", m.Absorption);
        e.num("smoothness", m.Smoothness); e.num("reflectivity", m.Reflectivity); e.color("tint", m.Tint);
        e.num("absorption_scale", m.AbsorptionScale);
        e.close('}');
    }
    e.close(']');
    e.key("settings"); e.open('{');
    e.integer("width", sc.Settings.Width); e.integer("height", sc.Settings.Height);
    e.integer("samples_per_px", sc.Settings.SamplesPerPx); e.integer("max_depth", sc.Settings.MaxDepth);
    e.close('}');
    e.color("background", sc.Background);
    if (sc.SkyPtr) {
        e.key("sky"); e.open('{');
        e.str("type", sc.SkyPtr->Type); e.color("color", sc.SkyPtr->SkyColor); e.color("horizon", sc.SkyPtr->Horizon);
        e.color("zenith", sc.SkyPtr->Zenith);
        e.close('}');
    } else { e.key("sky"); e.out += "null"; }
    if (sc.FogPtr) {
        const Fog& f = *sc.FogPtr;
        e.key("fog"); e.open('{');
        e.num("density", f.Density); e.color("color", f.FogColor); e.num("scatter", f.Scatter); e.num("sigma_s", f.SigmaS);
        e.num("sigma_a", f.SigmaA); e.num("g", f.G); e.num("hetero_strength", f.HeteroStrength); e.num("noise_scale", f.NoiseScale);
        e.integer("noise_octaves", f.NoiseOctaves); e.boolean("affect_sky", f.AffectSky); e.boolean("gpu_volumetric", f.GPUVolumetric);
        e.close('}');
    }
    e.close('}');
    e.out += "\n";   // Encoder.Encode appends a newline
    return e.out;
}

void Save(const std::string& path, const Scene& sc) {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("create scene: open " + path + ": " + std::strerror(errno));
    std::string s = Marshal(sc);
    f.write(s.data(), (std::streamsize)s.size());
    if (!f) throw std::runtime_error("encode scene: write failed");
}

// ------------------------------------------------------------------ flattening (the host half of sceneToWorld)
static int32_t material_code(const std::string& t) {           // materials.go:33-54 switch
    if (t == MaterialMetal) return PTB_MAT_METAL;
    if (t == MaterialDielectric) return PTB_MAT_DIELECTRIC;
    if (t == MaterialEmissive) return PTB_MAT_EMISSIVE;
    if (t == MaterialMirror) return PTB_MAT_MIRROR;
    return PTB_MAT_LAMBERT;
}
static int32_t object_code(const std::string& t) {             // objects.go:237-266 switch
    if (t == ObjectSphere || t == ObjectSphereLight) return PTB_OBJ_SPHERE;
    if (t == ObjectPlane) return PTB_OBJ_PLANE;
    if (t == ObjectBox) return PTB_OBJ_BOX;
    if (t == ObjectMesh) return PTB_OBJ_MESH;                  // EXTENSION
    return -1;                                                 // dropped
}

uint64_t MeshData::NextGeneration() {
    static std::atomic<uint64_t> next{1};
    return next.fetch_add(1);
}

Flat Flatten(const Scene& sc) {
    Flat f;
    uint64_t gen = 0xcbf29ce484222325ull;
    auto fold = [&gen](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { gen ^= b[i]; gen *= 0x100000001b3ull; } };
    std::map<std::string, int> by_id;                          // objects.go:226-229: later duplicate id wins
    for (size_t i = 0; i < sc.Materials.size(); i++) {
        const Material& m = sc.Materials[i];
        by_id[m.ID] = (int)i;
        f.mat_type.push_back(material_code(m.Type));
        f.mat_albedo.insert(f.mat_albedo.end(), {m.Albedo.R, m.Albedo.G, m.Albedo.B});
        f.mat_rough.push_back(m.Rough);
        f.mat_ior.push_back(m.IOR);
        f.mat_emit.insert(f.mat_emit.end(), {m.Emit.R, m.Emit.G, m.Emit.B});
        f.mat_power.push_back(m.Power);
        f.mat_absorption.insert(f.mat_absorption.end(), {m.Absorption.R, m.Absorption.G, m.Absorption.B});
        f.mat_smoothness.push_back(m.Smoothness);
    }
    f.mesh_tri_begin.push_back(0);
    for (const Object& o : sc.Objects) {
        int32_t code = object_code(o.Type);
        if (code == PTB_OBJ_MESH && (!o.Mesh || o.Mesh->triangles.empty())) code = -1;   // a mesh object without triangles is dropped
        if (code == PTB_OBJ_MESH) {
            f.obj_mesh.push_back((int32_t)f.mesh_tri_begin.size() - 1);
            const double sx = o.Size.X != 0 ? o.Size.X : 1.0, sy = o.Size.Y != 0 ? o.Size.Y : 1.0, sz = o.Size.Z != 0 ? o.Size.Z : 1.0;
            const MeshData& m = *o.Mesh;
            {   // identity of this object's world-space triangles: the mesh data's generation and the placement
                const double place[6] = {o.Position.X, o.Position.Y, o.Position.Z, sx, sy, sz};
                fold(&m.generation, sizeof m.generation); fold(place, sizeof place);
            }
            for (uint32_t idx : m.triangles) {
                f.tri_vertices.push_back((float)(o.Position.X + sx * (double)m.vertices[3 * idx]));
                f.tri_vertices.push_back((float)(o.Position.Y + sy * (double)m.vertices[3 * idx + 1]));
                f.tri_vertices.push_back((float)(o.Position.Z + sz * (double)m.vertices[3 * idx + 2]));
            }
            f.mesh_tri_begin.push_back((int64_t)f.tri_vertices.size() / 9);
        } else {
            f.obj_mesh.push_back(-1);
        }
        f.obj_type.push_back(code);
        auto it = by_id.find(o.MaterialID);
        f.obj_mat.push_back(it == by_id.end() ? -1 : it->second);
        f.obj_pos.insert(f.obj_pos.end(), {o.Position.X, o.Position.Y, o.Position.Z});
        f.obj_size.insert(f.obj_size.end(), {o.Size.X, o.Size.Y, o.Size.Z});
    }
    f.mesh_generation = f.mesh_tri_begin.size() > 1 ? (gen | 1) : 0;     // non-zero whenever there is a mesh
    const Camera& c = sc.Cam;
    f.camera = ptb_camera{{c.Position.X, c.Position.Y, c.Position.Z}, {c.Target.X, c.Target.Y, c.Target.Z}, {c.Up.X, c.Up.Y, c.Up.Z},
                          c.FOV, c.Aperture, c.FocusDist, c.AspectRatio};
    // sky selection, renderer.go:56-92
    f.sky = ptb_sky{};
    if (sc.SkyPtr && sc.SkyPtr->Type == "gradient") {
        f.sky.kind = PTB_SKY_GRADIENT;
        const Sky& s = *sc.SkyPtr;
        f.sky.horizon[0] = s.Horizon.R; f.sky.horizon[1] = s.Horizon.G; f.sky.horizon[2] = s.Horizon.B;
        f.sky.zenith[0] = s.Zenith.R; f.sky.zenith[1] = s.Zenith.G; f.sky.zenith[2] = s.Zenith.B;
    } else {
        f.sky.kind = PTB_SKY_CONST;
        Color bg = (sc.SkyPtr && sc.SkyPtr->Type == "solid") ? sc.SkyPtr->SkyColor : sc.Background;
        f.sky.color[0] = bg.R; f.sky.color[1] = bg.G; f.sky.color[2] = bg.B;
    }
    return f;
}

// ---- EXTENSION: heightfield generator.  y(x,z) = amplitude * sum_o 0.5^o * vnoise(x*f*2^o, z*f*2^o) with vnoise =
// smoothstep-interpolated lattice values in [-1,1] from an integer hash of (ix, iz, seed, octave).  binary64 throughout.
static double lattice(int ix, int iz, uint32_t seed, int oct) {
    uint32_t h = (uint32_t)ix * 0x9E3779B1u ^ (uint32_t)iz * 0x85EBCA77u ^ (seed + (uint32_t)oct * 0xC2B2AE3Du);
    h ^= h >> 16; h *= 0x21f0aaadu; h ^= h >> 15; h *= 0x735a2d97u; h ^= h >> 15;
    return (double)h * (2.0 / 4294967296.0) - 1.0;
}
static double vnoise(double x, double z, uint32_t seed, int oct) {
    const double fx = std::floor(x), fz = std::floor(z);
    const int ix = (int)fx, iz = (int)fz;
    double tx = x - fx, tz = z - fz;
    tx = tx * tx * (3.0 - 2.0 * tx); tz = tz * tz * (3.0 - 2.0 * tz);
    const double a = lattice(ix, iz, seed, oct), b = lattice(ix + 1, iz, seed, oct);
    const double c = lattice(ix, iz + 1, seed, oct), d = lattice(ix + 1, iz + 1, seed, oct);
    return (a + (b - a) * tx) + ((c + (d - c) * tx) - (a + (b - a) * tx)) * tz;
}
void GenerateHeightfield(MeshData& m) {
    if (m.nx < 1 || m.nz < 1 || (long long)m.nx * m.nz > 50000000LL) throw std::runtime_error("decode scene: heightfield nx/nz out of range");
    const int oct = m.octaves > 0 ? m.octaves : 4;
    const double freq = m.frequency != 0 ? m.frequency : 4.0;
    m.vertices.resize((size_t)(m.nx + 1) * (m.nz + 1) * 3);
    for (int j = 0; j <= m.nz; j++)
        for (int i = 0; i <= m.nx; i++) {
            const double x = (double)i / m.nx - 0.5, z = (double)j / m.nz - 0.5;
            double y = 0, w = 1, f = freq;
            for (int o = 0; o < oct; o++) { y += w * vnoise((x + 0.5) * f, (z + 0.5) * f, m.seed, o); w *= 0.5; f *= 2; }
            float* v = &m.vertices[((size_t)j * (m.nx + 1) + i) * 3];
            v[0] = (float)x; v[1] = (float)(m.amplitude * y); v[2] = (float)z;
        }
    m.triangles.resize((size_t)m.nx * m.nz * 6);
    size_t k = 0;
    for (int j = 0; j < m.nz; j++)
        for (int i = 0; i < m.nx; i++) {
            const uint32_t a = (uint32_t)(j * (m.nx + 1) + i), b = a + 1, c = a + (uint32_t)(m.nx + 1), d = c + 1;
            m.triangles[k++] = a; m.triangles[k++] = c; m.triangles[k++] = b;      // counter-clockwise seen from +y
            m.triangles[k++] = b; m.triangles[k++] = c; m.triangles[k++] = d;
        }
}

ptb_scene Flat::view() const {
    ptb_scene s{};
    s.n_obj = (int32_t)obj_type.size();
    s.obj_type = obj_type.data(); s.obj_mat = obj_mat.data(); s.obj_pos = obj_pos.data(); s.obj_size = obj_size.data();
    s.n_mat = (int32_t)mat_type.size();
    s.mat_type = mat_type.data(); s.mat_albedo = mat_albedo.data(); s.mat_rough = mat_rough.data(); s.mat_ior = mat_ior.data();
    s.mat_emit = mat_emit.data(); s.mat_power = mat_power.data(); s.mat_absorption = mat_absorption.data();
    s.mat_smoothness = mat_smoothness.data();
    s.camera = camera; s.sky = sky;
    s.n_mesh = (int32_t)mesh_tri_begin.size() - 1;
    s.obj_mesh = obj_mesh.data(); s.mesh_tri_begin = mesh_tri_begin.data(); s.tri_vertices = tri_vertices.data();
    s.mesh_generation = mesh_generation;
    return s;
}

}  // namespace scene
