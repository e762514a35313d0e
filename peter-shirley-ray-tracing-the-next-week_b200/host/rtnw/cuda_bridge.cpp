// The four virtuals of the reference API that do arithmetic (hitable::hit, material::scatter, material::emitted,
// texture::value; PSC/hitable.h:34, PSC/material.h:54-57, PSC/texture.h:13) served by the GPU library: each call
// flattens the object it is made on (cached per object), uploads the tables and issues a one-element
// rtnw_trace / rtnw_scatter / rtnw_eval_texture.  This keeps user code written against the reference API working
// (a single-ray query costs a kernel launch; bulk work belongs in rtnw_render / rtnw_trace).  Linked only into
// programs that also link librtnw.so; librtnw_host.so itself has no CUDA dependency.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

#include "rtnw.h"
#include "rtnw/device_bridge.hpp"
#include "rtnw/flatten.hpp"

namespace rtnw {
namespace {

struct resident {
    flat_scene flat;
    rtnw_scene* dev = nullptr;
};

rtnw_ctx* g_ctx = nullptr;
uint64_t g_seed = 1;
uint32_t g_calls = 0;  // every scatter call gets its own path stream (pixel slot = call number)
std::map<const void*, resident*> g_cache;

[[noreturn]] void die(const char* what) {
    std::fprintf(stderr, "rtnw bridge: %s: %s\n", what, rtnw_last_error());
    std::abort();
}

resident* resident_for(const void* key, const hitable* world) {
    auto it = g_cache.find(key);
    if (it != g_cache.end()) return it->second;
    resident* r = new resident();
    if (flatten(world, r->flat) != RTNW_OK) {
        std::fprintf(stderr, "rtnw bridge: cannot flatten: %s\n", r->flat.error.c_str());
        std::abort();
    }
    const rtnw_scene_desc d = r->flat.desc();
    if (rtnw_scene_upload(g_ctx, &d, &r->dev) != RTNW_OK) die("scene upload");
    g_cache[key] = r;
    return r;
}

// a one-sphere world carrying the material / texture under test
resident* probe_for_material(const material* m) {
    if (g_cache.count(m)) return g_cache[m];
    hitable** l = new hitable*[1];
    l[0] = new sphere(vec3(0, 0, 0), 1, const_cast<material*>(m));
    return resident_for(m, new hitable_list(l, 1));
}
resident* probe_for_texture(const texture* t) {
    if (g_cache.count(t)) return g_cache[t];
    hitable** l = new hitable*[1];
    l[0] = new sphere(vec3(0, 0, 0), 1, new lambertian(const_cast<texture*>(t)));
    return resident_for(t, new hitable_list(l, 1));
}

rtnw_ray to_c(const ray& r, uint32_t key) {
    rtnw_ray q;
    for (int c = 0; c < 3; ++c) { q.origin[c] = r.A.e[c]; q.direction[c] = r.B.e[c]; }
    q.time = r._time;
    q.key = key;
    return q;
}

bool bridge_hit(const hitable* h, const ray& r, float t_min, float t_max, hit_record& rec) {
    resident* s = resident_for(h, h);
    const rtnw_ray q = to_c(r, g_calls++);
    rtnw_hit out;
    if (rtnw_trace(g_ctx, s->dev, &q, 1, t_min, t_max, 0, g_seed, &out) != RTNW_OK) die("rtnw_trace");
    if (out.prim_id < 0) return false;
    rec.t = out.t; rec.u = out.u; rec.v = out.v;
    rec.p = vec3(out.p[0], out.p[1], out.p[2]);
    rec.normal = vec3(out.normal[0], out.normal[1], out.normal[2]);
    rec.mat_ptr = const_cast<material*>(s->flat.material_objects[out.mat_id]);
    return true;
}

void scatter_call(const material* m, const ray& r_in, const hit_record& rec, rtnw_ray& sc, float att[3], float em[3], int32_t& flag) {
    resident* s = probe_for_material(m);
    const rtnw_ray q = to_c(r_in, 0);
    rtnw_hit h;
    std::memset(&h, 0, sizeof h);
    h.t = rec.t; h.u = rec.u; h.v = rec.v;
    for (int c = 0; c < 3; ++c) { h.p[c] = rec.p.e[c]; h.normal[c] = rec.normal.e[c]; }
    h.mat_id = 0;
    for (size_t i = 0; i < s->flat.material_objects.size(); ++i)
        if (s->flat.material_objects[i] == m) h.mat_id = (int32_t)i;
    // rtnw_scatter keys the stream by the element index; vary the seed per call so repeated calls draw fresh numbers
    if (rtnw_scatter(g_ctx, s->dev, &q, &h, 1, g_seed + 0x9E3779B97F4A7C15ull * (uint64_t)(g_calls++), &sc, att, em, &flag) != RTNW_OK)
        die("rtnw_scatter");
}

bool bridge_scatter(const material* m, const ray& r_in, const hit_record& rec, vec3& attenuation, ray& scattered) {
    rtnw_ray sc; float att[3], em[3]; int32_t flag;
    scatter_call(m, r_in, rec, sc, att, em, flag);
    attenuation = vec3(att[0], att[1], att[2]);
    scattered = ray(vec3(sc.origin[0], sc.origin[1], sc.origin[2]), vec3(sc.direction[0], sc.direction[1], sc.direction[2]), sc.time);
    return flag != 0;
}

vec3 bridge_emitted(const material* m, float u, float v, const vec3& p) {
    hit_record rec;
    rec.t = 1; rec.u = u; rec.v = v; rec.p = p; rec.normal = vec3(0, 1, 0); rec.mat_ptr = const_cast<material*>(m);
    rtnw_ray sc; float att[3], em[3]; int32_t flag;
    scatter_call(m, ray(vec3(0, 0, 0), vec3(0, -1, 0), 0), rec, sc, att, em, flag);
    return vec3(em[0], em[1], em[2]);
}

vec3 bridge_value(const texture* t, float u, float v, const vec3& p) {
    resident* s = probe_for_texture(t);
    int32_t id = 0;
    for (size_t i = 0; i < s->flat.texture_objects.size(); ++i)
        if (s->flat.texture_objects[i] == t) id = (int32_t)i;
    const float uvp[5] = {u, v, p.e[0], p.e[1], p.e[2]};
    float rgb[3];
    if (rtnw_eval_texture(g_ctx, s->dev, id, uvp, 1, rgb) != RTNW_OK) die("rtnw_eval_texture");
    return vec3(rgb[0], rgb[1], rgb[2]);
}

ray bridge_get_ray(const camera* cam, float s, float t) {
    rtnw_camera c;
    to_c_camera(*cam, c);
    const float st[2] = {s, t};
    rtnw_ray out;
    if (rtnw_camera_get_rays(g_ctx, &c, st, 1, g_seed, g_calls++, &out) != RTNW_OK) die("rtnw_camera_get_rays");
    return ray(vec3(out.origin[0], out.origin[1], out.origin[2]), vec3(out.direction[0], out.direction[1], out.direction[2]), out.time);
}

}  // namespace

// Route the reference API's virtuals to GPU `device`.  Returns an rtnw_status.
int install_cuda_bridge(int device, unsigned long long seed) {
    if (!g_ctx && rtnw_ctx_create(device, &g_ctx) != RTNW_OK) return RTNW_ERR_CUDA;
    g_seed = seed;
    device_bridge b;
    b.hit = bridge_hit;
    b.scatter = bridge_scatter;
    b.emitted = bridge_emitted;
    b.value = bridge_value;
    b.get_ray = bridge_get_ray;
    set_device_bridge(b);
    return RTNW_OK;
}

rtnw_ctx* bridge_context() { return g_ctx; }

// The bridge keeps the flattened, uploaded form of every object a call was made on, keyed by the object's address.  An object
// that is mutated (or freed and its address reused) after its first call must be dropped from that cache:
void bridge_invalidate(const void* object) {
    auto it = g_cache.find(object);
    if (it == g_cache.end()) return;
    if (it->second->dev) rtnw_scene_free(g_ctx, it->second->dev);
    delete it->second;
    g_cache.erase(it);
}

// drop every cached scene and the bridge's context (device memory is returned); the virtuals abort again until the next install
void bridge_release() {
    for (auto& kv : g_cache) {
        if (kv.second->dev) rtnw_scene_free(g_ctx, kv.second->dev);
        delete kv.second;
    }
    g_cache.clear();
    if (g_ctx) rtnw_ctx_destroy(g_ctx);
    g_ctx = nullptr;
    set_device_bridge(device_bridge());
}

}  // namespace rtnw
