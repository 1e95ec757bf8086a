// Native restatement of the reference's scene ingest: SceneLoader.loadFromJSON and its helpers
// (reference js/scene-loader.js:20-284), the Camera constructor (js/camera.js:8-36), the Plane / light
// constructors' normalisation (js/geometry.js:52, js/lights.js:38) and TriangleMesh's triangle filtering
// (js/geometry.js:206-231).  Every default and skip rule is kept because object / triangle IDs depend on them.
// JavaScript value semantics that matter here: `a || b` (falsy: undefined, null, false, 0, NaN, ""), `!== undefined`,
// Array.isArray, and arithmetic coercion of null/booleans.  Values outside docs/scene_format.md (strings where
// numbers belong) become NaN.
#include <cmath>
#include <cstring>
#include <limits>
#include "brt_host.hpp"
#include "json.hpp"

namespace brt {

using brtjson::Value;
static const double kNaN = std::numeric_limits<double>::quiet_NaN();
static const double kPi = 3.141592653589793;

static bool truthy(const Value* v) {
    if (!v) return false;
    switch (v->kind) {
    case brtjson::NUL: return false;
    case brtjson::BOOL: return v->b;
    case brtjson::NUM: return !(v->num == 0 || v->num != v->num);
    case brtjson::STR: return !v->str.empty();
    default: return true;                       // [] and {} are truthy
    }
}
static double num(const Value* v) {
    if (!v) return kNaN;                         // undefined
    switch (v->kind) {
    case brtjson::NUL: return 0.0;
    case brtjson::BOOL: return v->b ? 1.0 : 0.0;
    case brtjson::NUM: return v->num;
    default: return kNaN;
    }
}
struct LoadError { std::string msg; };           // stands in for the TypeError caught at ray-tracer.js:330-333

static std::string lower(const Value* v, const char* what) {
    if (!v || !v->is_string()) throw LoadError{ std::string(what) + ".toLowerCase is not a function" };
    std::string s = v->str;
    for (auto& c : s) if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    return s;
}
static void parse_vec3(const Value* v, double out[3]) {          // scene-loader.js:268-273
    if (v && v->is_array() && v->arr.size() >= 3) { out[0] = num(&v->arr[0]); out[1] = num(&v->arr[1]); out[2] = num(&v->arr[2]); }
    else out[0] = out[1] = out[2] = 0.0;
}
static void set3(double d[3], double x, double y, double z) { d[0] = x; d[1] = y; d[2] = z; }
static double len3(const double v[3]) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static void normalize3(double v[3]) {                            // math.js:18
    double l = len3(v);
    if (l > 0) { v[0] /= l; v[1] /= l; v[2] /= l; } else set3(v, 0, 0, 0);
}

static int create_material(HostScene& sc, const Value* m) {      // scene-loader.js:143-173
    brt_material out{};
    out.type = BRT_MAT_LAMBERTIAN; set3(out.color, 0.8, 0.8, 0.8); out.param = 0;
    if (truthy(m) && truthy(m->get("type"))) {
        std::string t = lower(m->get("type"), "material.type");
        if (t == "lambertian") { parse_vec3(m->get("color"), out.color); }
        else if (t == "metal") {
            out.type = BRT_MAT_METAL; parse_vec3(m->get("color"), out.color);
            double r = m->get("roughness") ? num(m->get("roughness")) : 0.0;
            out.param = (r != r) ? r : std::fmin(r, 1.0);         // Math.min(roughness, 1) (materials.js:33)
        } else if (t == "dielectric") {
            out.type = BRT_MAT_DIELECTRIC; set3(out.color, 1, 1, 1);
            out.param = m->get("ior") ? num(m->get("ior")) : 1.5;
        } else if (t == "emissive") {
            out.type = BRT_MAT_EMISSIVE; parse_vec3(m->get("color"), out.color);
            out.param = m->get("intensity") ? num(m->get("intensity")) : 1.0;
        }                                                         // unknown type: default lambertian (:169-171)
    }
    sc.materials.push_back(out);
    return (int)sc.materials.size() - 1;
}

static void fetch_vertex(const std::vector<double>& verts, size_t nverts, double idx, double out[3]) {
    // vertices[idx] for a non-element index reads `undefined` -> _ensureVec3 -> (0,0,0)  (geometry.js:240-246)
    if (!(idx >= 0) || idx != std::floor(idx) || idx >= (double)nverts) { set3(out, 0, 0, 0); return; }
    size_t k = (size_t)idx;
    set3(out, verts[3 * k], verts[3 * k + 1], verts[3 * k + 2]);
}

// Binary container (brt_scene_load_binary): mesh arrays may live in a blob after the JSON header instead of as JSON text.
struct Blob { const unsigned char* p = nullptr; size_t n = 0; };
static thread_local Blob g_blob;

// {"offset": bytes from blob start, "count": elements} -> pointer, or throws
static const unsigned char* blob_ref(const Value* ref, size_t elemBytes, size_t& count) {
    if (!ref || !ref->is_object()) throw LoadError{ "binary mesh reference must be an object {offset, count}" };
    double off = num(ref->get("offset")), cnt = num(ref->get("count"));
    if (!(off >= 0) || !(cnt >= 0) || off != std::floor(off) || cnt != std::floor(cnt)) throw LoadError{ "bad binary mesh reference" };
    if (!g_blob.p || off + cnt * (double)elemBytes > (double)g_blob.n) throw LoadError{ "binary mesh reference outside the blob" };
    count = (size_t)cnt;
    return g_blob.p + (size_t)off;
}

static void create_object(HostScene& sc, const Value& o) {       // scene-loader.js:90-137
    if (!o.is_object()) {
        if (o.kind == brtjson::NUL) throw LoadError{ "cannot read properties of null (reading 'type')" };
        return;                                                   // numbers/strings/arrays: .type is undefined -> skipped
    }
    if (!truthy(o.get("type"))) return;
    const Value* md = o.get("material");
    int mat = create_material(sc, truthy(md) ? md : nullptr);
    std::string t = lower(o.get("type"), "type");
    brt_object ob{};
    ob.material = mat;
    if (t == "sphere") {
        ob.type = BRT_OBJ_SPHERE; parse_vec3(o.get("center"), ob.a);
        ob.b[0] = truthy(o.get("radius")) ? num(o.get("radius")) : 1.0;       // `objData.radius || 1.0` (:101)
    } else if (t == "plane") {
        ob.type = BRT_OBJ_PLANE; parse_vec3(o.get("point"), ob.a); parse_vec3(o.get("normal"), ob.b);
        normalize3(ob.b);                                                     // geometry.js:52
    } else if (t == "box") {
        ob.type = BRT_OBJ_BOX; parse_vec3(o.get("min"), ob.a); parse_vec3(o.get("max"), ob.b);
    } else if (t == "triangle") {
        ob.type = BRT_OBJ_TRIANGLE; parse_vec3(o.get("v0"), ob.a); parse_vec3(o.get("v1"), ob.b); parse_vec3(o.get("v2"), ob.c);
    } else if (t == "mesh" && o.get("vertices_bin")) {
        // same TriangleMesh rules (geometry.js:206-231) over float64 vertex triples and uint32 indices stored in the blob
        size_t nvd = 0, ni = 0;
        const unsigned char* vb = blob_ref(o.get("vertices_bin"), sizeof(double), nvd);
        const unsigned char* ib = blob_ref(o.get("indices_bin"), sizeof(uint32_t), ni);
        const size_t nv = nvd / 3;
        ob.type = BRT_OBJ_MESH; ob.first_tri = (int64_t)(sc.meshTris.size() / 9); ob.tri_count = 0;
        sc.meshTris.reserve(sc.meshTris.size() + 3 * ni);
        for (size_t i = 0; i + 2 < ni; i += 3) {
            uint32_t ix[3];
            memcpy(ix, ib + 4 * i, 12);
            if (ix[0] >= nv || ix[1] >= nv || ix[2] >= nv) continue;             // geometry.js:216-219
            double tri[9];
            for (int k = 0; k < 3; k++) memcpy(tri + 3 * k, vb + 24 * (size_t)ix[k], 24);
            sc.meshTris.insert(sc.meshTris.end(), tri, tri + 9);
            ob.tri_count++;
        }
    } else if (t == "mesh") {
        const Value* vs = o.get("vertices"); const Value* is = o.get("indices");
        if (!truthy(vs) || !truthy(is)) { sc.materials.pop_back(); return; }  // :120-123
        if (!vs->is_array()) throw LoadError{ "vertices.map is not a function" };
        ob.type = BRT_OBJ_MESH; ob.first_tri = (int64_t)(sc.meshTris.size() / 9); ob.tri_count = 0;
        if (is->is_array()) {                                                 // else: geometry.js:199-202 -> empty mesh
            size_t nv = vs->arr.size();
            std::vector<double> verts(3 * nv);
            for (size_t k = 0; k < nv; k++) parse_vec3(&vs->arr[k], &verts[3 * k]);
            size_t ni = is->arr.size();
            for (size_t i = 0; i + 2 < ni; i += 3) {                          // geometry.js:206-210 (incomplete tail skipped)
                double ix[3];
                for (int k = 0; k < 3; k++) { const Value& e = is->arr[i + k]; ix[k] = e.kind == brtjson::NUM ? e.num : kNaN; }
                if (ix[0] >= (double)nv || ix[1] >= (double)nv || ix[2] >= (double)nv) continue;   // :216-219
                double tri[9];
                for (int k = 0; k < 3; k++) fetch_vertex(verts, nv, ix[k], tri + 3 * k);
                sc.meshTris.insert(sc.meshTris.end(), tri, tri + 9);
                ob.tri_count++;
            }
        }
    } else { sc.materials.pop_back(); return; }                               // unknown type (:133-135)
    sc.objects.push_back(ob);
}

static void create_light(HostScene& sc, const Value& l) {        // scene-loader.js:179-200
    if (!truthy(&l) || !truthy(l.get("type"))) return;
    brt_light out{};
    const Value* col = l.get("color");
    if (truthy(col)) parse_vec3(col, out.color); else set3(out.color, 1, 1, 1);
    out.intensity = l.get("intensity") ? num(l.get("intensity")) : 1.0;
    std::string t = lower(l.get("type"), "light.type");
    if (t == "point") { out.type = BRT_LIGHT_POINT; parse_vec3(l.get("position"), out.v); }
    else if (t == "directional") { out.type = BRT_LIGHT_DIRECTIONAL; parse_vec3(l.get("direction"), out.v); normalize3(out.v); }   // lights.js:38
    else return;
    sc.lights.push_back(out);
}

void derive_camera(brt_camera& c) {                              // camera.js:8-36
    double theta = c.vfov * kPi / 180;
    double h = std::tan(theta / 2);
    double viewportHeight = 2.0 * h;
    double viewportWidth = c.aspect * viewportHeight;
    double w[3] = { c.look_from[0] - c.look_at[0], c.look_from[1] - c.look_at[1], c.look_from[2] - c.look_at[2] };
    normalize3(w);
    double u[3] = { c.vup[1] * w[2] - c.vup[2] * w[1], c.vup[2] * w[0] - c.vup[0] * w[2], c.vup[0] * w[1] - c.vup[1] * w[0] };
    normalize3(u);
    double v[3] = { w[1] * u[2] - w[2] * u[1], w[2] * u[0] - w[0] * u[2], w[0] * u[1] - w[1] * u[0] };
    for (int k = 0; k < 3; k++) { c.w[k] = w[k]; c.u[k] = u[k]; c.v[k] = v[k]; c.origin[k] = c.look_from[k]; }
    if (c.type == BRT_CAM_PERSPECTIVE) {                          // :25
        for (int k = 0; k < 3; k++) {
            c.horizontal[k] = u[k] * (viewportWidth * c.focus_dist);
            c.vertical[k] = v[k] * (viewportHeight * c.focus_dist);
            c.lower_left_corner[k] = c.origin[k] - c.horizontal[k] / 2 - c.vertical[k] / 2 - w[k] * c.focus_dist;
        }
    } else {
        for (int k = 0; k < 3; k++) {
            c.horizontal[k] = u[k] * viewportWidth;
            c.vertical[k] = v[k] * viewportHeight;
            c.lower_left_corner[k] = c.origin[k] - c.horizontal[k] / 2 - c.vertical[k] / 2;
        }
    }
    c.lens_radius = c.aperture / 2;
}

static void create_camera(const Value& cd, double aspect, brt_camera& c) {   // scene-loader.js:205-262
    static const Value none;
    const Value* v;
    double defPos[3] = { 0, 0, 5 }, defUp[3] = { 0, 1, 0 };
    v = cd.get("position"); if (truthy(v)) parse_vec3(v, c.look_from); else set3(c.look_from, defPos[0], defPos[1], defPos[2]);
    v = cd.get("lookAt"); if (truthy(v)) parse_vec3(v, c.look_at); else set3(c.look_at, 0, 0, 0);
    v = cd.get("up"); if (truthy(v)) parse_vec3(v, c.vup); else set3(c.vup, defUp[0], defUp[1], defUp[2]);
    c.vfov = cd.get("fov") ? num(cd.get("fov")) : 45.0;
    c.aperture = cd.get("aperture") ? num(cd.get("aperture")) : 0.0;
    double d[3] = { c.look_from[0] - c.look_at[0], c.look_from[1] - c.look_at[1], c.look_from[2] - c.look_at[2] };
    double dist = len3(d);
    if (dist < 1.0) {                                                          // :213-224
        normalize3(d);
        for (int k = 0; k < 3; k++) c.look_at[k] = c.look_from[k] + (d[k] * -1) * 100;
    }
    if (cd.get("focusDist")) c.focus_dist = num(cd.get("focusDist"));
    else {                                                                     // :228-233
        double t[3] = { c.look_from[0] - c.look_at[0], c.look_from[1] - c.look_at[1], c.look_from[2] - c.look_at[2] };
        c.focus_dist = len3(t);
    }
    const Value* ty = cd.get("type");                                          // `camData.type || 'perspective'`, NOT lower-cased (:235)
    if (!truthy(ty)) c.type = BRT_CAM_PERSPECTIVE;
    else if (ty->is_string() && ty->str == "perspective") c.type = BRT_CAM_PERSPECTIVE;
    else if (ty->is_string() && ty->str == "orthographic") c.type = BRT_CAM_ORTHOGRAPHIC;
    else c.type = BRT_CAM_OTHER;
    c.aspect = truthy(cd.get("aspect")) ? num(cd.get("aspect")) : aspect;      // :236
    c.use_derived = 0;
    derive_camera(c);
}

int load_scene_json(const char* utf8, size_t len, int fallbackW, int fallbackH, HostScene& scene, HostBackground& bg, brt_camera& cam,
                    bool& hasCamera, int& outW, int& outH, std::string& err) {
    Value root;
    if (!brtjson::parse(utf8, len, root, err)) return BRT_E_PARSE;
    try {
        if (!root.is_object()) throw LoadError{ "scene root is not an object" };
        double width = fallbackW, height = fallbackH;
        outW = outH = 0;
        const Value* cd = root.get("camera");
        if (truthy(cd) && truthy(cd->get("resolution"))) {                     // scene-loader.js:25-34
            const Value* r = cd->get("resolution");
            if (!r->is_array() || r->arr.size() < 2) throw LoadError{ "camera.resolution must be [width, height]" };
            width = num(&r->arr[0]); height = num(&r->arr[1]);
            if (!(width >= 1) || !(height >= 1) || width > 65536 || height > 65536) throw LoadError{ "camera.resolution out of range" };
            outW = (int)width; outH = (int)height;
        }
        HostScene sc;
        HostBackground nb = bg;                                                // keeps the Perlin table of this ctx
        nb.kind = BRT_BG_GRADIENT; nb.intensity = 1.0; set3(nb.color, 0.1, 0.1, 0.1);   // new World() (world.js:12-13)
        const Value* b = root.get("background");
        if (truthy(b)) {                                                       // scene-loader.js:38-56
            const Value* t = b->get("type");
            if (t && t->is_string()) {
                if (t->str == "solid") nb.kind = BRT_BG_SOLID;
                else if (t->str == "hdri") nb.kind = BRT_BG_HDRI;
                else if (t->str == "procedural_sky") nb.kind = BRT_BG_PROCEDURAL_SKY;
            }
            // Deviation D1: scene-loader.js:43,45 bind the factory instead of calling it, so the raw
            // loader yields NaN -> black for solid/hdri; we implement the intended behaviour (ray-tracer.js:573-576).
            const Value* col = b->get("color");
            if (truthy(col)) parse_vec3(col, nb.color);
            if (b->get("intensity")) nb.intensity = num(b->get("intensity"));
        }
        const Value* objs = root.get("objects");
        if (truthy(objs) && objs->is_array()) for (const Value& o : objs->arr) create_object(sc, o);
        const Value* ls = root.get("lights");
        if (truthy(ls) && ls->is_array()) for (const Value& l : ls->arr) create_light(sc, l);
        hasCamera = false;
        brt_camera nc{};
        if (truthy(cd)) { create_camera(*cd, width / height, nc); hasCamera = true; }
        scene = std::move(sc);
        bg = nb;
        if (hasCamera) cam = nc;
        return BRT_OK;
    } catch (const LoadError& e) {
        err = e.msg;
        return BRT_E_PARSE;
    }
}

// Container: "BRTSCN01" | u64 json_len (LE) | json (docs/scene_format.md, meshes may carry vertices_bin / indices_bin
// = {offset, count} instead of vertices / indices) | zero padding to 8 bytes | blob (float64 vertex triples, uint32 indices).
int load_scene_binary(const unsigned char* bytes, size_t len, int fallbackW, int fallbackH, HostScene& scene, HostBackground& bg,
                      brt_camera& cam, bool& hasCamera, int& outW, int& outH, std::string& err) {
    if (len < 16 || memcmp(bytes, "BRTSCN01", 8) != 0) { err = "not a BRTSCN01 container"; return BRT_E_PARSE; }
    uint64_t jl = 0;
    memcpy(&jl, bytes + 8, 8);
    if (jl > len - 16) { err = "truncated BRTSCN01 container"; return BRT_E_PARSE; }
    size_t blobAt = 16 + (size_t)jl;
    blobAt = (blobAt + 7) & ~(size_t)7;
    g_blob.p = blobAt <= len ? bytes + blobAt : nullptr;
    g_blob.n = blobAt <= len ? len - blobAt : 0;
    int rc = load_scene_json((const char*)bytes + 16, (size_t)jl, fallbackW, fallbackH, scene, bg, cam, hasCamera, outW, outH, err);
    g_blob = Blob{};
    return rc;
}

}  // namespace brt
