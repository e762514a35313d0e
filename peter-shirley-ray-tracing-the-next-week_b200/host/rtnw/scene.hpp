// rtnw/scene.hpp — host-side scene-construction API, source compatible with the builders of the reference
// (`Peter-Shirley-Project Code/main.cpp:49-230`, PSC/ below): same global class names and constructor
// signatures, raw-`new` ownership, objects freely shared.  Nothing here intersects rays on the CPU: the object
// graph is a *description* that rtnw::flatten() lowers into the SoA tables of include/rtnw.h, and the virtual
// hit()/scatter()/value() entry points of the reference API are served by the GPU library (see device_bridge).
//
// Numerical contract: every constructor computes exactly what the reference constructor computes, with the same
// float/double promotions, because the values land in device tables that must be bit-identical to the
// reference's object state (camera basis PSC/camera.h:21-39, rotate_y sin/cos and box PSC/hitable.h:98-126,
// rect padding PSC/aarect.h:16-42, moving-sphere centres PSC/sphere.h:81-90, bvh_node topology PSC/bvh.h:97-121).
#ifndef RTNW_SCENE_HPP_
#define RTNW_SCENE_HPP_

#include <cmath>
#include <cstdlib>
#include <cfloat>
#include <iosfwd>
#include <iostream>
#include <vector>

#ifndef MAXFLOAT
#define MAXFLOAT FLT_MAX
#endif

// ------------------------------------------------------------------------------------------------ math types
// PSC/vec3.h:12-145.  Three floats; every operation is component-wise float arithmetic.
class vec3 {
public:
    float e[3];
    vec3() {}
    vec3(float a, float b, float c) : e{a, b, c} {}
    float x() const { return e[0]; }
    float y() const { return e[1]; }
    float z() const { return e[2]; }
    float r() const { return e[0]; }
    float g() const { return e[1]; }
    float b() const { return e[2]; }
    const vec3& operator+() const { return *this; }
    vec3 operator-() const { return vec3(-e[0], -e[1], -e[2]); }
    float operator[](int i) const { return e[i]; }
    float& operator[](int i) { return e[i]; }
#define RTNW_V3_COMPOUND(OP)                                   \
    vec3& operator OP(const vec3& o) {                         \
        for (int i = 0; i < 3; ++i) e[i] OP o.e[i];            \
        return *this;                                          \
    }
    RTNW_V3_COMPOUND(+=) RTNW_V3_COMPOUND(-=) RTNW_V3_COMPOUND(*=) RTNW_V3_COMPOUND(/=)
#undef RTNW_V3_COMPOUND
    vec3& operator*=(const float s) { for (float& c : e) c *= s; return *this; }
    // the reference scales by a float reciprocal computed in double (PSC/vec3.h:134-141)
    vec3& operator/=(const float s) { const float k = 1.0 / s; for (float& c : e) c *= k; return *this; }
    float squared_length() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    float length() const { return std::sqrt(squared_length()); }
    void make_unit_vector() { const float k = 1.0 / std::sqrt(squared_length()); for (float& c : e) c *= k; }
};
#define RTNW_V3_BINARY(OP)                                                                  \
    inline vec3 operator OP(const vec3& a, const vec3& b) {                                 \
        return vec3(a.e[0] OP b.e[0], a.e[1] OP b.e[1], a.e[2] OP b.e[2]);                  \
    }
RTNW_V3_BINARY(+) RTNW_V3_BINARY(-) RTNW_V3_BINARY(*) RTNW_V3_BINARY(/)
#undef RTNW_V3_BINARY
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.e[0], s * a.e[1], s * a.e[2]); }
inline vec3 operator*(const vec3& a, float s) { return s * a; }
inline vec3 operator/(vec3 a, float s) { return vec3(a.e[0] / s, a.e[1] / s, a.e[2] / s); }
inline float dot(const vec3& a, const vec3& b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.e[1] * b.e[2] - a.e[2] * b.e[1], -(a.e[0] * b.e[2] - a.e[2] * b.e[0]), a.e[0] * b.e[1] - a.e[1] * b.e[0]);
}
inline vec3 unit_vector(vec3 a) { return a / a.length(); }
std::istream& operator>>(std::istream& is, vec3& t);
std::ostream& operator<<(std::ostream& os, const vec3& t);

// PSC/ray.h:11-25
class ray {
public:
    vec3 A, B;
    float _time;
    ray() {}
    ray(const vec3& a, const vec3& b, float ti = 0.0) : A(a), B(b), _time(ti) {}
    vec3 origin() const { return A; }
    vec3 direction() const { return B; }
    float time() const { return _time; }
    vec3 point_at_parameter(float t) const { return A + t * B; }
};

// PSC/aabb.h:21-62.  Only min()/max()/surrounding_box feed the flattener; the slab test itself runs on the device.
class aabb {
public:
    vec3 _min, _max;
    aabb() {}
    aabb(const vec3& a, const vec3& b) : _min(a), _max(b) {}
    vec3 min() const { return _min; }
    vec3 max() const { return _max; }
};
inline aabb surrounding_box(aabb b0, aabb b1) {
    vec3 lo(std::fmin(b0._min.e[0], b1._min.e[0]), std::fmin(b0._min.e[1], b1._min.e[1]), std::fmin(b0._min.e[2], b1._min.e[2]));
    vec3 hi(std::fmax(b0._max.e[0], b1._max.e[0]), std::fmax(b0._max.e[1], b1._max.e[1]), std::fmax(b0._max.e[2], b1._max.e[2]));
    return aabb(lo, hi);
}

// ------------------------------------------------------------------------------------------------ appearance
namespace rtnw {
enum class tex_kind { constant, checker, noise, image, readme_noise, user };
enum class mat_kind { lambertian, metal, dielectric, diffuse_light, isotropic, user };
enum class geo_kind { sphere, moving_sphere, rect_xy, rect_xz, rect_yz, box, flip, translate, rotate_y, list, bvh, medium, user };
}  // namespace rtnw

struct hit_record;
class material;

// PSC/texture.h:11-14.  value() is evaluated on the GPU (rtnw_eval_texture); see device_bridge.hpp.
class texture {
public:
    virtual ~texture() {}
    virtual vec3 value(float u, float v, const vec3& p) const;
    virtual rtnw::tex_kind rtnw_kind() const { return rtnw::tex_kind::user; }
};
class constant_texture : public texture {  // PSC/texture.h:16-28
public:
    vec3 color;
    constant_texture() {}
    constant_texture(vec3 c) : color(c) {}
    rtnw::tex_kind rtnw_kind() const override { return rtnw::tex_kind::constant; }
};
class checker_texture : public texture {  // PSC/texture.h:30-45 (first ctor argument is `even`)
public:
    texture* odd;
    texture* even;
    checker_texture() : odd(nullptr), even(nullptr) {}
    checker_texture(texture* t0, texture* t1) : odd(t1), even(t0) {}
    rtnw::tex_kind rtnw_kind() const override { return rtnw::tex_kind::checker; }
};
// PSC/perlin.h:41-111: the four tables are process-wide, drawn from drand48 in the order ranvec, perm_x, perm_y,
// perm_z.  The reference draws them during static initialisation; rtnw::perlin_tables() does so on first use (or
// on regenerate()) so the host controls where in the drand48 stream they fall.
class perlin {
public:
    static vec3* ranvec;
    static int* perm_x;
    static int* perm_y;
    static int* perm_z;
    static void regenerate();  // 768 + 3*255 draws
    // the tables of the reference's Chapter 4 drafts (README.md:536-570): 256 floats `ranfloat` (kept in ranvec[i].x), then the
    // three permutations — 256 + 3*255 draws
    static void regenerate_readme();
};
// The three intermediate noise textures of README.md:516-630 (hash only / trilinear / Hermite-smoothed trilinear), value =
// (1,1,1) * noise(p): regression scenes for the shipped Chapter04 images.  Needs perlin::regenerate_readme() tables.
class readme_noise_texture : public texture {
public:
    int variant;  // 1, 2, 3
    explicit readme_noise_texture(int v) : variant(v) {}
    rtnw::tex_kind rtnw_kind() const override { return rtnw::tex_kind::readme_noise; }
};
class noise_texture : public texture {  // PSC/texture.h:47-59
public:
    perlin noise;
    float scale;
    noise_texture() : scale(1.0f) {}  // the reference leaves `scale` uninitialised; 1 is our defined value
    noise_texture(float sc) : scale(sc) {}
    rtnw::tex_kind rtnw_kind() const override { return rtnw::tex_kind::noise; }
};
class image_texture : public texture {  // PSC/surface_texture.h:10-30, tightly packed RGB8
public:
    unsigned char* data;
    int nx, ny;
    bool bilinear;  // option of this framework (RTNW_TEXF_BILINEAR); the reference's lookup is nearest-texel
    image_texture() : data(nullptr), nx(0), ny(0), bilinear(false) {}
    image_texture(unsigned char* pixels, int A, int B, bool blend = false) : data(pixels), nx(A), ny(B), bilinear(blend) {}
    rtnw::tex_kind rtnw_kind() const override { return rtnw::tex_kind::image; }
};

// PSC/material.h:50-58.  scatter()/emitted() are evaluated on the GPU (rtnw_scatter).
class material {
public:
    virtual ~material() {}
    virtual bool scatter(const ray& r_in, const hit_record& rec, vec3& attenuation, ray& scattered) const;
    virtual vec3 emitted(float u, float v, const vec3& p) const;
    virtual rtnw::mat_kind rtnw_kind() const { return rtnw::mat_kind::user; }
};
class lambertian : public material {  // PSC/material.h:61-72
public:
    texture* albedo;
    lambertian(texture* a) : albedo(a) {}
    rtnw::mat_kind rtnw_kind() const override { return rtnw::mat_kind::lambertian; }
};
class metal : public material {  // PSC/material.h:74-85, fuzz clamped at construction
public:
    vec3 albedo;
    float fuzz;
    metal(const vec3& a, float f) : albedo(a), fuzz(f < 1 ? f : 1) {}
    rtnw::mat_kind rtnw_kind() const override { return rtnw::mat_kind::metal; }
};
class dielectric : public material {  // PSC/material.h:87-123
public:
    float ref_idx;
    dielectric(float ri) : ref_idx(ri) {}
    rtnw::mat_kind rtnw_kind() const override { return rtnw::mat_kind::dielectric; }
};
class diffuse_light : public material {  // PSC/material.h:126-139
public:
    texture* emit;
    diffuse_light(texture* a) : emit(a) {}
    rtnw::mat_kind rtnw_kind() const override { return rtnw::mat_kind::diffuse_light; }
};
class isotropic : public material {  // PSC/material.h:142-151
public:
    texture* albedo;
    isotropic(texture* a) : albedo(a) {}
    rtnw::mat_kind rtnw_kind() const override { return rtnw::mat_kind::isotropic; }
};

// ------------------------------------------------------------------------------------------------ geometry
struct hit_record {  // PSC/hitable.h:21-29
    float t, u, v;
    vec3 p, normal;
    material* mat_ptr;
};

// PSC/hitable.h:31-36.  hit() is a single-ray query served by the GPU (rtnw_trace); bounding_box() is host
// arithmetic because it defines the BVH the device traverses.
class hitable {
public:
    virtual ~hitable() {}
    virtual bool hit(const ray& r, float t_min, float t_max, hit_record& rec) const;
    virtual bool bounding_box(float t0, float t1, aabb& box) const = 0;
    virtual rtnw::geo_kind rtnw_kind() const { return rtnw::geo_kind::user; }
};

class sphere : public hitable {  // PSC/sphere.h:10-58
public:
    vec3 center;
    float radius;
    material* mat_ptr;
    sphere() : radius(0), mat_ptr(nullptr) {}
    sphere(vec3 cen, float r, material* m) : center(cen), radius(r), mat_ptr(m) {}
    bool bounding_box(float, float, aabb& box) const override {
        const vec3 rr(radius, radius, radius);
        box = aabb(center - rr, center + rr);
        return true;
    }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::sphere; }
};

class moving_sphere : public hitable {  // PSC/sphere.h:61-118
public:
    vec3 center0, center1;
    float time0, time1, radius;
    material* mat_ptr;
    moving_sphere() : time0(0), time1(1), radius(0), mat_ptr(nullptr) {}
    moving_sphere(vec3 cen0, vec3 cen1, float t0, float t1, float r, material* m)
        : center0(cen0), center1(cen1), time0(t0), time1(t1), radius(r), mat_ptr(m) {}
    vec3 center(float time) const { return center0 + ((time - time0) / (time1 - time0)) * (center1 - center0); }
    bool bounding_box(float t0, float t1, aabb& box) const override {
        const vec3 rr(radius, radius, radius);
        box = surrounding_box(aabb(center(t0) - rr, center(t0) + rr), aabb(center(t1) - rr, center(t1) + rr));
        return true;
    }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::moving_sphere; }
};

namespace rtnw {
// One axis-aligned rectangle type for the three reference classes (PSC/aarect.h:11-100): plane axis N, extent
// axes A < B.  The slab of the bounding box is padded by 0.0001 in double, then rounded (PSC/aarect.h:17,29,41).
template <int N>
class aarect : public hitable {
public:
    material* mp;
    float a0, a1, b0, b1, k;
    aarect() : mp(nullptr), a0(0), a1(0), b0(0), b1(0), k(0) {}
    aarect(float _a0, float _a1, float _b0, float _b1, float _k, material* mat) : mp(mat), a0(_a0), a1(_a1), b0(_b0), b1(_b1), k(_k) {}
    bool bounding_box(float, float, aabb& box) const override {
        const float lo = k - 0.0001, hi = k + 0.0001;
        if (N == 2) box = aabb(vec3(a0, b0, lo), vec3(a1, b1, hi));
        else if (N == 1) box = aabb(vec3(a0, lo, b0), vec3(a1, hi, b1));
        else box = aabb(vec3(lo, a0, b0), vec3(hi, a1, b1));
        return true;
    }
    geo_kind rtnw_kind() const override { return N == 2 ? geo_kind::rect_xy : (N == 1 ? geo_kind::rect_xz : geo_kind::rect_yz); }
};
}  // namespace rtnw
typedef rtnw::aarect<2> xy_rect;  // xy_rect(x0,x1,y0,y1,k,mat)
typedef rtnw::aarect<1> xz_rect;  // xz_rect(x0,x1,z0,z1,k,mat)
typedef rtnw::aarect<0> yz_rect;  // yz_rect(y0,y1,z0,z1,k,mat)

class flip_normals : public hitable {  // PSC/hitable.h:39-54
public:
    hitable* ptr;
    flip_normals(hitable* p) : ptr(p) {}
    bool bounding_box(float t0, float t1, aabb& box) const override { return ptr->bounding_box(t0, t1, box); }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::flip; }
};

class translate : public hitable {  // PSC/hitable.h:57-83
public:
    hitable* ptr;
    vec3 offset;
    translate(hitable* p, const vec3& displacement) : ptr(p), offset(displacement) {}
    bool bounding_box(float t0, float t1, aabb& box) const override {
        if (!ptr->bounding_box(t0, t1, box)) return false;
        box = aabb(box.min() + offset, box.max() + offset);
        return true;
    }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::translate; }
};

class rotate_y : public hitable {  // PSC/hitable.h:85-150
public:
    hitable* ptr;
    float sin_theta, cos_theta;
    bool hasbox;
    aabb bbox;
    rotate_y(hitable* p, float angle);
    bool bounding_box(float, float, aabb& box) const override { box = bbox; return hasbox; }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::rotate_y; }
};

class hitable_list : public hitable {  // PSC/hitable_list.h:10-50
public:
    hitable** list;
    int list_size;
    hitable_list() : list(nullptr), list_size(0) {}
    hitable_list(hitable** l, int n) : list(l), list_size(n) {}
    // true union of the children (the reference unions list[0] n times; documented deviation)
    bool bounding_box(float t0, float t1, aabb& box) const override;
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::list; }
};

class box : public hitable {  // PSC/box.h:11-38; list_ptr holds the six faces in the reference's order
public:
    vec3 pmin, pmax;
    hitable* list_ptr;
    material* mat_ptr;
    box() : list_ptr(nullptr), mat_ptr(nullptr) {}
    box(const vec3& p0, const vec3& p1, material* ptr);
    bool bounding_box(float, float, aabb& b) const override { b = aabb(pmin, pmax); return true; }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::box; }
};

class constant_medium : public hitable {  // PSC/constant_medium.h:14-50
public:
    hitable* boundary;
    float density;
    material* phase_function;
    constant_medium(hitable* b, float d, texture* a) : boundary(b), density(d), phase_function(new isotropic(a)) {}
    bool bounding_box(float t0, float t1, aabb& box) const override { return boundary->bounding_box(t0, t1, box); }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::medium; }
};

// PSC/bvh.h:11-121.  The constructor reproduces the reference build step for step — one drand48 per node for
// the axis, libc qsort on the children's bounding_box(0,0).min()[axis] with the reference's never-equal
// comparator, split at n/2 — so that the topology (and therefore the tie order among equal-t hits) is the
// reference's.  `l` is sorted in place, as in the reference.
class bvh_node : public hitable {
public:
    hitable* left;
    hitable* right;
    aabb box;
    float time0, time1;
    std::vector<hitable*> creation_order;  // children as handed in, before sorting (defines leaf ids); top node only
    bvh_node() : left(nullptr), right(nullptr), time0(0), time1(0) {}
    bvh_node(hitable** l, int n, float t0, float t1);
    bool bounding_box(float, float, aabb& b) const override { b = box; return true; }
    rtnw::geo_kind rtnw_kind() const override { return rtnw::geo_kind::bvh; }
private:
    struct inner_tag {};
    bvh_node(hitable** l, int n, float t0, float t1, inner_tag);
    void build(hitable** l, int n);
};

// PSC/camera.h:10-58.  get_ray() is evaluated on the GPU (rtnw_camera_rays); the constructor is host arithmetic.
class camera {
public:
    vec3 origin, u, v, w, horizontal, vertical, lower_left_corner;
    float len_radius, time0, time1;
    camera() {}
    camera(vec3 lookfrom, vec3 lookat, vec3 vup, float vfov, float aspect, float aperture, float focus_dist, float t0, float t1);
    ray get_ray(float s, float t) const;  // PSC/camera.h:41-47; evaluated on the GPU like the other arithmetic entry points (device_bridge)
};

#endif  // RTNW_SCENE_HPP_
