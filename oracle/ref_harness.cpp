// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// Compiles the UNMODIFIED reference renderer (`/root/reference/Peter-Shirley-Project Code/main.cpp` and the 15
// headers it includes, PSC/ below) into oracle/_ref/libref_oracle.so behind a small C ABI, so that tests can
// pin the oracle port (oracle/rtnw_oracle.c) and generate golden vectors, and bench.py can time the reference's
// own CPU implementation (`cpu_baseline.kind = "reference"`).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the result.
//
// How the reference is reached (no reference source is copied into the repo; oracle/Makefile copies the
// sources to a temp dir at build time, applies the two edits below with sed, and deletes the temp dir):
//   * single TU: `#define main ref_unused_main` + `#include "main.cpp"`;
//   * F2 patch: aabb.h:38-39 subtract r.origin() instead of r.direction() (without it bvh_node::hit finds
//     nothing) — every BVH result of this oracle is "reference + F2 patch";
//   * counters: `ref_cnt[k]++` at the top of aabb::hit / sphere::hit / moving_sphere::hit / x?_rect::hit /
//     constant_medium::hit, for the per-ray test counts the roofline uses;
//   * `#define drand48 ref_hook_drand48`: scene construction still gets glibc's drand48; while rendering with
//     rng_mode=1 the draws come from the framework's stream (Philox4x32-10 keyed by (seed; pixel, sample) seeding
//     the drand48 recurrence, media drawing keyed Philox numbers) so the
//     reference and the GPU consume the SAME random numbers and their images can be compared sample for sample.
//
// Leaf ids: hit_record has no primitive id, so every leaf handed to the list/BVH is wrapped in a
// `tagged` forwarder that swaps rec.mat_ptr for a per-leaf proxy material carrying (leaf id, box face).
#include <assert.h>
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cmath>
#include <fstream>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "rtnw.h"  // struct layouts only (rtnw_ray, rtnw_hit)

// ---------------------------------------------------------------------------------------------- RNG hook
namespace {

struct philox_state {
    uint32_t key[2];
    uint32_t pixel, sample;
    uint64_t x;          // state of the path's sequential stream (drand48 recurrence, seeded by Philox)
    int seeded;          // x holds the state of the current (pixel, sample)
    int depth;           // index of the current top-level closest-hit query in this path
    int leaf;            // >= 0 while inside a tagged leaf's hit() (medium free-flight draws are keyed by it)
    int mode;            // 0 = glibc drand48 (scene construction, native baseline), 1 = Philox streams
    uint64_t draws;
} G = {{0, 0}, 0, 0, 0, 0, 0, -1, 0, 0};

inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

inline double u01(uint32_t x) { return (double)((float)(x >> 8) * (1.0f / 16777216.0f)); }

double hook_stream_draw() {
    G.draws++;
    if (G.leaf >= 0) {  // keyed: (pixel, sample, depth, leaf) -> one number, independent of traversal order
        const uint32_t ctr[4] = {(uint32_t)G.leaf, 1u + (uint32_t)G.depth, G.sample, G.pixel};
        uint32_t out[4];
        philox4x32_10(ctr, G.key, out);
        return u01(out[0]);
    }
    // sequential stream of the path: Philox(0, 0, sample, pixel) seeds the drand48 recurrence
    if (!G.seeded) {
        const uint32_t ctr[4] = {0u, 0u, G.sample, G.pixel};
        uint32_t out[4];
        philox4x32_10(ctr, G.key, out);
        G.x = ((uint64_t)(out[1] & 0xffffu) << 32) | (uint64_t)out[0];
        G.seeded = 1;
    }
    G.x = (G.x * 0x5DEECE66DULL + 0xBULL) & 0xffffffffffffULL;
    return (double)((float)(uint32_t)(G.x >> 24) * (1.0f / 16777216.0f));
}

inline void begin_path(uint32_t pixel, uint32_t sample) {
    G.pixel = pixel;
    G.sample = sample;
    G.seeded = 0;
    G.depth = -1;
    G.leaf = -1;
}

}  // namespace

static inline double ref_hook_drand48() { return G.mode == 0 ? drand48() : hook_stream_draw(); }

long ref_cnt[8];  // 0 aabb, 1 sphere, 2 moving_sphere, 3 rect, 4 medium

#define drand48 ref_hook_drand48
#define main ref_unused_main
#include "main.cpp"  // the reference, from the temp copy on the include path
#undef main
#undef drand48

// ---------------------------------------------------------------------------------------------- tagging
namespace {

struct proxy_material : public material {
    material* real;
    int leaf, sub, kind;
    proxy_material(int l, int s, int k) : real(nullptr), leaf(l), sub(s), kind(k) {}
    virtual bool scatter(const ray& r_in, const hit_record& rec, vec3& attenuation, ray& scattered) const {
        return real->scatter(r_in, rec, attenuation, scattered);
    }
    virtual vec3 emitted(float u, float v, const vec3& p) const { return real->emitted(u, v, p); }
};

enum leaf_kind { LK_SPHERE = 0, LK_MOVING = 1, LK_RECT_XY = 2, LK_RECT_XZ = 3, LK_RECT_YZ = 4, LK_BOX = 5, LK_MEDIUM = 6, LK_OTHER = 7 };

struct tagged : public hitable {
    hitable* inner;
    proxy_material* proxy;
    tagged(hitable* h, int leaf, int sub, int kind) : inner(h), proxy(new proxy_material(leaf, sub, kind)) {}
    virtual bool hit(const ray& r, float t_min, float t_max, hit_record& rec) const {
        const int saved = G.leaf;
        G.leaf = proxy->leaf;
        const bool h = inner->hit(r, t_min, t_max, rec);
        G.leaf = saved;
        if (h) {
            proxy->real = rec.mat_ptr;
            rec.mat_ptr = proxy;
        }
        return h;
    }
    virtual bool bounding_box(float t0, float t1, aabb& b) const { return inner->bounding_box(t0, t1, b); }
};

struct counted_world : public hitable {  // counts rays = top-level closest-hit queries (PSC/main.cpp:27)
    hitable* inner;
    mutable uint64_t rays;
    explicit counted_world(hitable* h) : inner(h), rays(0) {}
    virtual bool hit(const ray& r, float t_min, float t_max, hit_record& rec) const {
        rays++;
        G.depth++;
        return inner->hit(r, t_min, t_max, rec);
    }
    virtual bool bounding_box(float t0, float t1, aabb& b) const { return inner->bounding_box(t0, t1, b); }
};

hitable* strip_wrappers(hitable* h) {
    for (;;) {
        if (translate* t = dynamic_cast<translate*>(h)) h = t->ptr;
        else if (rotate_y* r = dynamic_cast<rotate_y*>(h)) h = r->ptr;
        else if (flip_normals* f = dynamic_cast<flip_normals*>(h)) h = f->ptr;
        else return h;
    }
}

int classify(hitable* core) {
    if (dynamic_cast<moving_sphere*>(core)) return LK_MOVING;
    if (dynamic_cast<sphere*>(core)) return LK_SPHERE;
    if (dynamic_cast<xy_rect*>(core)) return LK_RECT_XY;
    if (dynamic_cast<xz_rect*>(core)) return LK_RECT_XZ;
    if (dynamic_cast<yz_rect*>(core)) return LK_RECT_YZ;
    if (dynamic_cast<box*>(core)) return LK_BOX;
    if (dynamic_cast<constant_medium*>(core)) return LK_MEDIUM;
    return LK_OTHER;
}

struct ref_scene_impl {
    hitable* world;               // what color() is called with (tagged or not)
    counted_world* counted;
    std::vector<hitable*> leaves; // untagged leaf slots in id order (for the dump)
    bool tagged_build;
    int next_id;
    ref_scene_impl() : world(NULL), counted(NULL), tagged_build(false), next_id(0) {}

    // Tag one leaf slot (a hitable handed to the list / BVH).  Boxes are tagged face by face so the proxy can
    // carry the face index; everything else is wrapped whole.
    hitable* tag(hitable* slot) {
        const int id = next_id++;
        leaves.push_back(slot);
        if (!tagged_build) return slot;
        hitable* core = strip_wrappers(slot);
        const int kind = classify(core);
        if (box* b = dynamic_cast<box*>(core)) {
            hitable_list* faces = static_cast<hitable_list*>(b->list_ptr);
            for (int k = 0; k < 6; ++k) faces->list[k] = new tagged(faces->list[k], id, k, kind);
            return slot;
        }
        return new tagged(slot, id, 0, kind);
    }
};

// ------------------------------------------------------------------------------------------ extra builders
// Config 1: TNW/Chapter01_Motion Blur.cpp:36-67 restated for the current API (lambertian takes a texture*;
// the snapshot no longer compiles).  Expression shapes are the snapshot's, so g++ draws in the
// same order it would for the snapshot.
hitable* h_random_scene_ch01() {
    int n = 500;
    hitable** list = new hitable*[n + 1];
    list[0] = new sphere(vec3(0, -700, 0), 700, new lambertian(new constant_texture(vec3(0.5, 0.5, 0.5))));
    int i = 1;
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            float choose_mat = ref_hook_drand48();
            vec3 center(a + 0.9 * ref_hook_drand48(), 0.2, b + 0.9 * ref_hook_drand48());
            if ((center - vec3(4, 0.2, 0)).length() > 0.9) {
                if (choose_mat < 0.8) {
                    list[i++] = new moving_sphere(center, center + vec3(0, 0.5 * ref_hook_drand48(), 0), 0.0, 1.0, 0.2,
                                                  new lambertian(new constant_texture(vec3(ref_hook_drand48() * ref_hook_drand48(),
                                                                                           ref_hook_drand48() * ref_hook_drand48(),
                                                                                           ref_hook_drand48() * ref_hook_drand48()))));
                } else if (choose_mat < 0.95) {
                    list[i++] = new sphere(center, 0.2,
                                           new metal(vec3(0.5 * (1 + ref_hook_drand48()), 0.5 * (1 + ref_hook_drand48()),
                                                          0.5 * (1 + ref_hook_drand48())), 0.5 * ref_hook_drand48()));
                } else {
                    list[i++] = new sphere(center, 0.2, new dielectric(1.5));
                }
            }
        }
    }
    list[i++] = new sphere(vec3(0, 1, 0), 1.0, new dielectric(2.5));
    list[i++] = new sphere(vec3(-4, 1, 0), 1.0, new lambertian(new constant_texture(vec3(0.4, 0.2, 0.1))));
    list[i++] = new sphere(vec3(4, 1, 0), 1.0, new metal(vec3(1, 1, 1), 0.0));
    return new hitable_list(list, i);
}

// Config 2: the shipped two_perlin_spheres() reads an uninitialised scale (F5).
hitable* h_two_perlin() {
    texture* checker = new checker_texture(new constant_texture(vec3(0.2, 0.3, 0.1)), new constant_texture(vec3(0.9, 0.9, 0.9)));
    hitable** list = new hitable*[2];
    list[0] = new sphere(vec3(0, -1000, 0), 1000, new lambertian(checker));
    list[1] = new sphere(vec3(0, 2, 0), 2, new lambertian(new noise_texture(4)));
    return new hitable_list(list, 2);
}

// The three intermediate noise functions of README.md Chapter 4, before perlin.h reached its shipped form (the sources of
// TNW/Chapter04_Perlin noise_noise{1,2 smoth,3 hermite cubic smoth}.ppm; SURVEY.md §4: "source not shipped").  Restated from the
// README excerpts: variant 1 = hash only (README.md:516-524), 2 = trilinear interpolation of the lattice values (README.md:599-611,
// i.e. the reference's own trilinear_interp, PSC/perlin.h:11-23), 3 = the same with Hermite-smoothed u,v,w (README.md:619-630).
// They index a table of 256 FLOATS (ranfloat[i] = drand48(), README.md:536-542), drawn before the three permutations.
// The texture is the book's at that stage: value = (1,1,1) * noise(p).
float* g_ranfloat = NULL;
struct readme_noise_texture : public texture {
    int variant;
    explicit readme_noise_texture(int v) : variant(v) {}
    float noise(const vec3& p) const {
        if (variant == 1) {
            int i = int(4 * p.x()) & 255;
            int j = int(4 * p.y()) & 255;
            int k = int(4 * p.z()) & 255;
            return g_ranfloat[perlin::perm_x[i] ^ perlin::perm_y[j] ^ perlin::perm_z[k]];
        }
        float u = p.x() - floor(p.x());
        float v = p.y() - floor(p.y());
        float w = p.z() - floor(p.z());
        if (variant == 3) {
            u = u * u * (3 - 2 * u);
            v = v * v * (3 - 2 * v);
            w = w * w * (3 - 2 * w);
        }
        int i = floor(p.x());
        int j = floor(p.y());
        int k = floor(p.z());
        float c[2][2][2];
        for (int di = 0; di < 2; di++)
            for (int dj = 0; dj < 2; dj++)
                for (int dk = 0; dk < 2; dk++)
                    c[di][dj][dk] = g_ranfloat[perlin::perm_x[(i + di) & 255] ^ perlin::perm_y[(j + dj) & 255] ^ perlin::perm_z[(k + dk) & 255]];
        return trilinear_interp(c, u, v, w);
    }
    virtual vec3 value(float u, float v, const vec3& p) const { return vec3(1, 1, 1) * noise(p); }
};

void regenerate_readme_tables() {  // README.md:536-570: ranfloat, then perm_x, perm_y, perm_z
    g_ranfloat = new float[256];
    for (int i = 0; i < 256; i++) g_ranfloat[i] = ref_hook_drand48();
    perlin::perm_x = perlin_generate_perm();
    perlin::perm_y = perlin_generate_perm();
    perlin::perm_z = perlin_generate_perm();
}

hitable* h_perlin_variant(int variant) {  // two_perlin_spheres() of that stage (README.md:588-596): both spheres carry the noise
    texture* pertext = new readme_noise_texture(variant);
    hitable** list = new hitable*[2];
    list[0] = new sphere(vec3(0, -1000, 0), 1000, new lambertian(pertext));
    list[1] = new sphere(vec3(0, 2, 0), 2, new lambertian(pertext));
    return new hitable_list(list, 2);
}

unsigned char* h_synthetic_earth(int& nx, int& ny) {  // same integer pattern as the host library's stand-in image
    nx = 1024;
    ny = 512;
    unsigned char* px = new unsigned char[(size_t)nx * ny * 3];
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            const unsigned h = (unsigned)(x / 32) * 2654435761u ^ (unsigned)(y / 32) * 40503u;
            const bool land = ((h >> 13) & 7u) < 3u;
            unsigned char* p = px + 3 * ((size_t)y * nx + x);
            p[0] = (unsigned char)(land ? 60 + ((x * 5 + y * 3) & 63) : 10 + (y & 31));
            p[1] = (unsigned char)(land ? 120 + ((x * 3 + y * 7) & 63) : 40 + ((x + y) & 63));
            p[2] = (unsigned char)(land ? 40 + ((x ^ y) & 31) : 140 + ((x * 2 + y) & 63));
        }
    return px;
}

hitable* h_earth() {  // PSC/main.cpp:87-97 with the synthetic image instead of stbi_load("picture.png")
    hitable** list = new hitable*[2];
    material* light = new diffuse_light(new constant_texture(vec3(7, 7, 7)));
    list[0] = new xz_rect(63, 483, 55, 482, 554, light);
    int nx, ny;
    unsigned char* tex_data = h_synthetic_earth(nx, ny);
    list[1] = new sphere(vec3(360, 250, 150), 100, new lambertian(new image_texture(tex_data, nx, ny)));
    return new hitable_list(list, 2);
}

// Config 5N: the builder of PSC/main.cpp:190-230 at north-star scale: nb = 32 floor boxes in
// a bvh_node, the 1000-sphere cluster in translate(rotate_y(bvh_node)), plus the image-textured sphere.
hitable* h_final_northstar(ref_scene_impl& S) {
    int nb = 32;
    hitable** list = new hitable*[30];
    hitable** boxlist = new hitable*[nb * nb];
    hitable** boxlist2 = new hitable*[1000];
    material* white = new lambertian(new constant_texture(vec3(0.73, 0.73, 0.73)));
    material* ground = new lambertian(new constant_texture(vec3(0.48, 0.83, 0.53)));
    int b = 0, l = 0;
    for (int i = 0; i < nb; i++) {
        for (int j = 0; j < nb; j++) {
            float w = 1000.0f / nb;
            float x0 = i * w;
            float z0 = j * w;
            float y0 = 0;
            float x1 = x0 + w;
            float y1 = 100 * (ref_hook_drand48() + 0.01);
            float z1 = z0 + w;
            boxlist[b++] = S.tag(new box(vec3(x0, y0, z0), vec3(x1, y1, z1), ground));
        }
    }
    list[l++] = new bvh_node(boxlist, b, 0, 1);
    material* light = new diffuse_light(new constant_texture(vec3(7, 7, 7)));
    list[l++] = S.tag(new xz_rect(123, 423, 147, 412, 554, light));
    vec3 center(400, 400, 200);
    list[l++] = S.tag(new moving_sphere(center, center + vec3(30, 0, 0), 0, 1, 50,
                                        new lambertian(new constant_texture(vec3(0.7, 0.3, 0.1)))));
    list[l++] = S.tag(new sphere(vec3(260, 150, 45), 50, new dielectric(1.5)));
    list[l++] = S.tag(new sphere(vec3(0, 150, 145), 50, new metal(vec3(0.8, 0.8, 0.9), 10.0)));
    hitable* boundary = new sphere(vec3(360, 150, 145), 70, new dielectric(1.5));
    list[l++] = S.tag(boundary);
    list[l++] = S.tag(new constant_medium(boundary, 0.2, new constant_texture(vec3(0.2, 0.4, 0.9))));
    boundary = new sphere(vec3(0, 0, 0), 5000, new dielectric(1.5));
    list[l++] = S.tag(new constant_medium(boundary, 0.0001, new constant_texture(vec3(1.0, 1.0, 1.0))));
    int nx, ny;
    unsigned char* tex_data = h_synthetic_earth(nx, ny);
    list[l++] = S.tag(new sphere(vec3(400, 200, 400), 100, new lambertian(new image_texture(tex_data, nx, ny))));
    texture* pertext = new noise_texture(0.1);
    list[l++] = S.tag(new sphere(vec3(220, 280, 300), 80, new lambertian(pertext)));
    int ns = 1000;
    for (int j = 0; j < ns; j++) {
        boxlist2[j] = S.tag(new sphere(vec3(165 * ref_hook_drand48(), 165 * ref_hook_drand48(), 165 * ref_hook_drand48()), 10, white));
    }
    list[l++] = new translate(new rotate_y(new bvh_node(boxlist2, ns, 0.0, 1.0), 15), vec3(-100, 270, 395));
    return new hitable_list(list, l);
}

// integrator with the chapter snapshots' variations: t_min, sky background
// (TNW/Chapter01_Motion Blur.cpp:14-33), emitted term on/off.  With (0.001, black, emit) it is PSC/main.cpp:25-46,
// and ref_render uses the reference's own color() for that setting.
vec3 color_variant(const ray& r, hitable* world, int depth, float t_min, int sky, int emit, int max_depth) {
    hit_record rec;
    if (world->hit(r, t_min, MAXFLOAT, rec)) {
        ray scattered;
        vec3 attenuation;
        vec3 emitted = emit ? rec.mat_ptr->emitted(rec.u, rec.v, rec.p) : vec3(0, 0, 0);
        if (depth < max_depth && rec.mat_ptr->scatter(r, rec, attenuation, scattered)) {
            if (emit) return emitted + attenuation * color_variant(scattered, world, depth + 1, t_min, sky, emit, max_depth);
            return attenuation * color_variant(scattered, world, depth + 1, t_min, sky, emit, max_depth);
        }
        return emitted;
    }
    if (sky) {
        vec3 unit_direction = unit_vector(r.direction());
        float t = 0.5 * (unit_direction.y() + 1.0);
        return (1.0 - t) * vec3(1.0, 1.0, 1.0) + t * vec3(0.5, 0.7, 1.0);
    }
    return vec3(0, 0, 0);
}

struct quiet_cout {  // final() prints every box (PSC/main.cpp:206,228)
    quiet_cout() { std::cout.setstate(std::ios_base::failbit); }
    ~quiet_cout() { std::cout.clear(); }
};

void regenerate_perlin_tables() {  // the reference's static initialisers, PSC/perlin.h:108-111, in the same order
    perlin::ranvec = perlin_generate();
    perlin::perm_x = perlin_generate_perm();
    perlin::perm_y = perlin_generate_perm();
    perlin::perm_z = perlin_generate_perm();
}

}  // namespace

// ---------------------------------------------------------------------------------------------- C ABI
extern "C" {

typedef struct ref_scene_impl ref_scene;

int ref_abi_version(void) { return RTNW_ABI_VERSION; }

// name as in rtnw_host_scene_build ("final", "cornell_box+bvh", ...); tagged != 0 wraps leaves for ids
ref_scene* ref_scene_build(const char* name_c, int tagged_build) {
    std::string name(name_c);
    bool wrap = false;
    const size_t plus = name.find("+bvh");
    if (plus != std::string::npos) { wrap = true; name.erase(plus); }
    quiet_cout quiet;
    G.mode = 0;
    srand48(0x1234ABCD);  // glibc's never-seeded state
    if (name.compare(0, 8, "perlin_v") == 0) regenerate_readme_tables(); else
    regenerate_perlin_tables();
    ref_scene_impl* S = new ref_scene_impl();
    S->tagged_build = tagged_build != 0;
    hitable* w = NULL;
    bool pre_tagged = false;
    if (name == "ch01_random") w = h_random_scene_ch01();
    else if (name == "two_perlin") w = h_two_perlin();
    else if (name == "cornell_box") w = cornell_box();
    else if (name == "cornell_smoke") w = cornell_smoke();
    else if (name == "final") w = final();
    else if (name == "simple_light") w = simple_light();
    else if (name == "two_spheres") w = two_spheres();
    else if (name == "earth") w = h_earth();
    else if (name == "random_scene") w = random_scene();
    else if (name == "test") w = test();
    else if (name == "final_northstar") { w = h_final_northstar(*S); pre_tagged = true; }
    else if (name == "perlin_v1" || name == "perlin_v2" || name == "perlin_v3") w = h_perlin_variant(name[8] - '0');
    else { delete S; return NULL; }
    hitable_list* flat = static_cast<hitable_list*>(w);
    if (!pre_tagged)
        for (int i = 0; i < flat->list_size; ++i) flat->list[i] = S->tag(flat->list[i]);
    if (wrap) w = new bvh_node(flat->list, flat->list_size, 0, 1);
    S->counted = new counted_world(w);
    S->world = S->counted;
    return S;
}

int ref_scene_leaf_count(const ref_scene* S) { return (int)S->leaves.size(); }

// 24 floats per leaf: [0] kind, [1..9] geometry, [10] n_wrappers, [11..18] two wrappers (type,a,b,c),
// [19] material kind, [20..23] material params.  Order = leaf id.
int ref_scene_dump(const ref_scene* S, float* out, int max_leaves) {
    const int n = (int)S->leaves.size() < max_leaves ? (int)S->leaves.size() : max_leaves;
    for (int id = 0; id < n; ++id) {
        float* o = out + 24 * id;
        for (int k = 0; k < 24; ++k) o[k] = 0;
        hitable* h = S->leaves[id];
        int nw = 0;
        int flips = 0;
        for (;;) {
            if (translate* t = dynamic_cast<translate*>(h)) {
                if (nw < 2) { o[11 + 4 * nw] = 1; o[12 + 4 * nw] = t->offset.x(); o[13 + 4 * nw] = t->offset.y(); o[14 + 4 * nw] = t->offset.z(); }
                nw++; h = t->ptr;
            } else if (rotate_y* r = dynamic_cast<rotate_y*>(h)) {
                if (nw < 2) { o[11 + 4 * nw] = 2; o[12 + 4 * nw] = r->sin_theta; o[13 + 4 * nw] = r->cos_theta; }
                nw++; h = r->ptr;
            } else if (flip_normals* f = dynamic_cast<flip_normals*>(h)) {
                flips++; h = f->ptr;
            } else break;
        }
        o[10] = (float)(nw + 100 * (flips & 1));
        material* m = NULL;
        o[0] = (float)classify(h);
        if (moving_sphere* s = dynamic_cast<moving_sphere*>(h)) {
            o[1] = s->center0.x(); o[2] = s->center0.y(); o[3] = s->center0.z(); o[4] = s->radius; o[5] = s->time0; o[6] = s->time1;
            o[7] = s->center1.x(); o[8] = s->center1.y(); o[9] = s->center1.z(); m = s->mat_ptr;
        } else if (sphere* s = dynamic_cast<sphere*>(h)) {
            o[1] = s->center.x(); o[2] = s->center.y(); o[3] = s->center.z(); o[4] = s->radius; m = s->mat_ptr;
        } else if (xy_rect* r = dynamic_cast<xy_rect*>(h)) {
            o[1] = r->x0; o[2] = r->x1; o[3] = r->y0; o[4] = r->y1; o[5] = r->k; m = r->mp;
        } else if (xz_rect* r = dynamic_cast<xz_rect*>(h)) {
            o[1] = r->x0; o[2] = r->x1; o[3] = r->z0; o[4] = r->z1; o[5] = r->k; m = r->mp;
        } else if (yz_rect* r = dynamic_cast<yz_rect*>(h)) {
            o[1] = r->y0; o[2] = r->y1; o[3] = r->z0; o[4] = r->z1; o[5] = r->k; m = r->mp;
        } else if (box* b = dynamic_cast<box*>(h)) {
            o[1] = b->pmin.x(); o[2] = b->pmin.y(); o[3] = b->pmin.z(); o[4] = b->pmax.x(); o[5] = b->pmax.y(); o[6] = b->pmax.z();
            hitable* f0 = static_cast<hitable_list*>(b->list_ptr)->list[0];
            if (tagged* tg = dynamic_cast<tagged*>(f0)) f0 = tg->inner;
            m = static_cast<xy_rect*>(f0)->mp;
        } else if (constant_medium* cm = dynamic_cast<constant_medium*>(h)) {
            o[1] = cm->density; m = cm->phase_function;
        }
        if (lambertian* lm = dynamic_cast<lambertian*>(m)) {
            o[19] = 0;
            if (constant_texture* ct = dynamic_cast<constant_texture*>(lm->albedo)) { o[20] = ct->color.x(); o[21] = ct->color.y(); o[22] = ct->color.z(); }
            else if (noise_texture* nt = dynamic_cast<noise_texture*>(lm->albedo)) { o[20] = -2; o[21] = nt->scale; }
            else if (dynamic_cast<checker_texture*>(lm->albedo)) o[20] = -1;
            else if (dynamic_cast<image_texture*>(lm->albedo)) o[20] = -3;
        } else if (metal* mm = dynamic_cast<metal*>(m)) {
            o[19] = 1; o[20] = mm->albedo.x(); o[21] = mm->albedo.y(); o[22] = mm->albedo.z(); o[23] = mm->fuzz;
        } else if (dielectric* dm = dynamic_cast<dielectric*>(m)) {
            o[19] = 2; o[23] = dm->ref_idx;
        } else if (diffuse_light* dl = dynamic_cast<diffuse_light*>(m)) {
            o[19] = 3;
            if (constant_texture* ct = dynamic_cast<constant_texture*>(dl->emit)) { o[20] = ct->color.x(); o[21] = ct->color.y(); o[22] = ct->color.z(); }
        } else if (isotropic* im = dynamic_cast<isotropic*>(m)) {
            o[19] = 4;
            if (constant_texture* ct = dynamic_cast<constant_texture*>(im->albedo)) { o[20] = ct->color.x(); o[21] = ct->color.y(); o[22] = ct->color.z(); }
        }
    }
    return n;
}

void ref_ranfloat_table(float* out256) {
    for (int i = 0; i < 256; ++i) out256[i] = g_ranfloat ? g_ranfloat[i] : 0.f;
}

void ref_perlin_tables(float* ranvec768, int* px, int* py, int* pz) {
    for (int i = 0; i < 256; ++i) {
        ranvec768[3 * i] = perlin::ranvec[i].x();
        ranvec768[3 * i + 1] = perlin::ranvec[i].y();
        ranvec768[3 * i + 2] = perlin::ranvec[i].z();
        px[i] = perlin::perm_x[i]; py[i] = perlin::perm_y[i]; pz[i] = perlin::perm_z[i];
    }
}

// one world->hit(r, t_min, t_max, rec) per ray (PSC/main.cpp:27); media draw keyed numbers (pixel slot = ray.key)
void ref_trace(ref_scene* S, const rtnw_ray* rays, size_t n, float t_min, float t_max, uint64_t seed, rtnw_hit* out) {
    G.mode = 1;
    G.key[0] = (uint32_t)seed;
    G.key[1] = (uint32_t)(seed >> 32);
    for (size_t i = 0; i < n; ++i) {
        const rtnw_ray& q = rays[i];
        begin_path(q.key, 0);
        ray r(vec3(q.origin[0], q.origin[1], q.origin[2]), vec3(q.direction[0], q.direction[1], q.direction[2]), q.time);
        hit_record rec;
        rec.t = 0; rec.u = 0; rec.v = 0; rec.p = vec3(0, 0, 0); rec.normal = vec3(0, 0, 0); rec.mat_ptr = NULL;
        rtnw_hit& h = out[i];
        memset(&h, 0, sizeof h);
        h.prim_id = -1;
        h.mat_id = -1;
        if (S->world->hit(r, t_min, t_max, rec)) {
            const proxy_material* pm = dynamic_cast<const proxy_material*>(rec.mat_ptr);
            h.prim_id = pm ? pm->leaf : -2;
            h.sub_id = pm ? pm->sub : 0;
            h.t = rec.t;
            for (int c = 0; c < 3; ++c) { h.p[c] = rec.p[c]; h.normal[c] = rec.normal[c]; }
            const bool has_uv = pm && pm->kind != LK_MOVING && pm->kind != LK_MEDIUM;  // the reference leaves u,v unwritten there (F5)
            h.u = has_uv ? rec.u : 0.0f;
            h.v = has_uv ? rec.v : 0.0f;
        }
    }
    G.mode = 0;
}

// The sample loop, PSC/main.cpp:299-313, for samples s = sample_begin + k*sample_stride.  accum = nx*ny*3 float
// sums, index (j*nx+i)*3+c.  rng_mode 0: glibc drand48 as shipped (srand48(seed) first unless seed == 0);
// rng_mode 1: framework Philox stream keyed by (seed, pixel, sample).  stats: [0] paths [1] rays [2] draws
// [3..7] aabb/sphere/moving/rect/medium test counts [8] seconds (CPU time of the loop only).
void ref_render(ref_scene* S, const float* lookfrom, const float* lookat, float vfov, float aperture, float focus_dist,
                float time0, float time1, int nx, int ny, int sample_begin, int sample_count, int sample_stride,
                int max_depth, float t_min, int sky, int emit, int denan, int rng_mode, uint64_t seed,
                float* accum, double* stats) {
    camera cam(vec3(lookfrom[0], lookfrom[1], lookfrom[2]), vec3(lookat[0], lookat[1], lookat[2]), vec3(0, 1, 0), vfov,
               float(nx) / float(ny), aperture, focus_dist, time0, time1);
    const bool stock_color = (t_min == 0.001f && !sky && emit && max_depth == 50);
    G.mode = rng_mode;
    G.key[0] = (uint32_t)seed;
    G.key[1] = (uint32_t)(seed >> 32);
    G.draws = 0;
    if (rng_mode == 0 && seed != 0) srand48((long)seed);
    for (int k = 0; k < 8; ++k) ref_cnt[k] = 0;
    S->counted->rays = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int j = ny - 1; j >= 0; j--) {
        for (int i = 0; i < nx; i++) {
            vec3 col(0, 0, 0);
            for (int k = 0; k < sample_count; k++) {
                const int s = sample_begin + k * sample_stride;
                begin_path((uint32_t)(j * nx + i), (uint32_t)s);
                float u = float(i + ref_hook_drand48()) / float(nx);
                float v = float(j + ref_hook_drand48()) / float(ny);
                ray r = cam.get_ray(u, v);
                vec3 temp = stock_color ? color(r, S->world, 0) : color_variant(r, S->world, 0, t_min, sky, emit, max_depth);
                if (denan) temp = de_nan(temp);
                col += temp;
            }
            accum[3 * (j * nx + i) + 0] = col[0];
            accum[3 * (j * nx + i) + 1] = col[1];
            accum[3 * (j * nx + i) + 2] = col[2];
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    G.mode = 0;
    if (stats) {
        stats[0] = (double)nx * ny * sample_count;
        stats[1] = (double)S->counted->rays;
        stats[2] = (double)G.draws;
        for (int k = 0; k < 5; ++k) stats[3 + k] = (double)ref_cnt[k];
        stats[8] = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    }
}

// The epilogue of the sample loop, PSC/main.cpp:315-325, compiled from the reference's OWN lines: oracle/Makefile cuts
// main.cpp:315-320 (mean, sqrt gamma, int(255.99*c)) and :322-325 (clamp to 255) into the two include files below (and
// checks their text), so this function is the reference's arithmetic on a buffer of sums.  out = nx*ny*3 ints, top row first.
void ref_epilogue(const float* sums, int nx, int ny, int ns, int clamp255, int* out) {
    size_t o = 0;
    for (int j = ny - 1; j >= 0; j--) {
        for (int i = 0; i < nx; i++) {
            const float* src = sums + 3 * ((size_t)j * nx + i);
            vec3 col(src[0], src[1], src[2]);
#include "ref_epilogue_a.inc"
            if (clamp255) {
#include "ref_epilogue_b.inc"
            }
            out[o++] = ir; out[o++] = ig; out[o++] = ib;
        }
    }
}

// camera::get_ray with the jitter of PSC/main.cpp:305-306, for n (i, j, s) triples
void ref_camera_rays(const float* lookfrom, const float* lookat, float vfov, float aperture, float focus_dist, float time0,
                     float time1, int nx, int ny, const int* ij, const int* sample, size_t n, uint64_t seed, rtnw_ray* out) {
    camera cam(vec3(lookfrom[0], lookfrom[1], lookfrom[2]), vec3(lookat[0], lookat[1], lookat[2]), vec3(0, 1, 0), vfov,
               float(nx) / float(ny), aperture, focus_dist, time0, time1);
    G.mode = 1;
    G.key[0] = (uint32_t)seed;
    G.key[1] = (uint32_t)(seed >> 32);
    for (size_t q = 0; q < n; ++q) {
        const int i = ij[2 * q], j = ij[2 * q + 1];
        begin_path((uint32_t)(j * nx + i), (uint32_t)sample[q]);
        float u = float(i + ref_hook_drand48()) / float(nx);
        float v = float(j + ref_hook_drand48()) / float(ny);
        ray r = cam.get_ray(u, v);
        for (int c = 0; c < 3; ++c) { out[q].origin[c] = r.origin()[c]; out[q].direction[c] = r.direction()[c]; }
        out[q].time = r.time();
        out[q].key = (uint32_t)(j * nx + i);
    }
    G.mode = 0;
}

// texture::value on freshly built reference textures. which: 0 constant(c0..c2), 1 checker of two constants
// (even = c0..c2, odd = c3..c5), 2 noise_texture(c0), 3 image_texture(synthetic earth).  uvp = n x 5.
void ref_eval_texture(int which, const float* c, const float* uvp, size_t n, float* rgb) {
    texture* t = NULL;
    if (which == 0) t = new constant_texture(vec3(c[0], c[1], c[2]));
    else if (which == 1) t = new checker_texture(new constant_texture(vec3(c[0], c[1], c[2])), new constant_texture(vec3(c[3], c[4], c[5])));
    else if (which == 2) t = new noise_texture(c[0]);
    else if (which >= 4 && which <= 6) t = new readme_noise_texture(which - 3);  // needs the tables of a perlin_v* scene build
    else { int nx, ny; unsigned char* d = h_synthetic_earth(nx, ny); t = new image_texture(d, nx, ny); }
    for (size_t i = 0; i < n; ++i) {
        const float* q = uvp + 5 * i;
        vec3 val = t->value(q[0], q[1], vec3(q[2], q[3], q[4]));
        rgb[3 * i] = val[0]; rgb[3 * i + 1] = val[1]; rgb[3 * i + 2] = val[2];
    }
}

// perlin::noise (which = 0) / perlin::turb (which = 1) with the current tables
void ref_eval_perlin(int which, const float* xyz, size_t n, float* out) {
    perlin p;
    for (size_t i = 0; i < n; ++i) {
        vec3 q(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        out[i] = which == 0 ? p.noise(q) : p.turb(q);
    }
}

// material::emitted + material::scatter.  mat = n x 8 floats {kind, tex_kind(0 const / 2 noise), r, g, b, fuzz_or_ri, scale, 0};
// the Philox sequential stream (seed, pixel = i, sample 0) supplies the draws.
void ref_scatter(const float* mat, const rtnw_ray* rays_in, const rtnw_hit* hits, size_t n, uint64_t seed,
                 rtnw_ray* out_scattered, float* out_atten, float* out_emitted, int* out_flag) {
    G.mode = 1;
    G.key[0] = (uint32_t)seed;
    G.key[1] = (uint32_t)(seed >> 32);
    for (size_t i = 0; i < n; ++i) {
        const float* m = mat + 8 * i;
        texture* tex = (int)m[1] == 2 ? (texture*)new noise_texture(m[6]) : (texture*)new constant_texture(vec3(m[2], m[3], m[4]));
        material* mp = NULL;
        switch ((int)m[0]) {
            case 0: mp = new lambertian(tex); break;
            case 1: mp = new metal(vec3(m[2], m[3], m[4]), m[5]); break;
            case 2: mp = new dielectric(m[5]); break;
            case 3: mp = new diffuse_light(tex); break;
            default: mp = new isotropic(tex); break;
        }
        begin_path((uint32_t)i, 0);
        const rtnw_ray& q = rays_in[i];
        ray r(vec3(q.origin[0], q.origin[1], q.origin[2]), vec3(q.direction[0], q.direction[1], q.direction[2]), q.time);
        hit_record rec;
        rec.t = hits[i].t; rec.u = hits[i].u; rec.v = hits[i].v;
        rec.p = vec3(hits[i].p[0], hits[i].p[1], hits[i].p[2]);
        rec.normal = vec3(hits[i].normal[0], hits[i].normal[1], hits[i].normal[2]);
        rec.mat_ptr = mp;
        vec3 em = mp->emitted(rec.u, rec.v, rec.p);
        vec3 att(0, 0, 0);
        ray sc(vec3(0, 0, 0), vec3(0, 0, 0), 0);
        const bool ok = mp->scatter(r, rec, att, sc);
        out_flag[i] = ok ? 1 : 0;
        for (int k = 0; k < 3; ++k) {
            out_emitted[3 * i + k] = em[k];
            out_atten[3 * i + k] = ok ? att[k] : 0.0f;
            out_scattered[i].origin[k] = ok ? sc.origin()[k] : 0.0f;
            out_scattered[i].direction[k] = ok ? sc.direction()[k] : 0.0f;
        }
        out_scattered[i].time = ok ? sc.time() : 0.0f;
        out_scattered[i].key = (uint32_t)i;  // mp/tex leak like everything in the reference (no virtual destructors)
    }
    G.mode = 0;
}

}  // extern "C"
