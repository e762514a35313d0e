// Host-side constructors of the scene API (see scene.hpp for the contract).  Each one reproduces the value the
// reference constructor leaves in the object, promotion for promotion, and nothing else runs on the CPU.
#include "rtnw/scene.hpp"
#include "rtnw/device_bridge.hpp"

#include <cstdio>
#include <istream>
#include <ostream>

std::istream& operator>>(std::istream& is, vec3& t) { return is >> t.e[0] >> t.e[1] >> t.e[2]; }
std::ostream& operator<<(std::ostream& os, const vec3& t) { return os << t.e[0] << " " << t.e[1] << " " << t.e[2]; }

// ---- API entry points that the reference evaluates on the CPU and we evaluate on the GPU ----------------------
namespace rtnw {
static device_bridge g_bridge;
void set_device_bridge(const device_bridge& b) { g_bridge = b; }
const device_bridge& get_device_bridge() { return g_bridge; }
[[noreturn]] static void no_bridge(const char* what) {
    std::fprintf(stderr,
                 "rtnw: %s is evaluated on the GPU; link librtnw.so and call rtnw::install_cuda_bridge() "
                 "(there is no CPU implementation of the path)\n", what);
    std::abort();
}
}  // namespace rtnw

bool hitable::hit(const ray& r, float t_min, float t_max, hit_record& rec) const {
    if (!rtnw::g_bridge.hit) rtnw::no_bridge("hitable::hit");
    return rtnw::g_bridge.hit(this, r, t_min, t_max, rec);
}
bool material::scatter(const ray& r_in, const hit_record& rec, vec3& attenuation, ray& scattered) const {
    if (!rtnw::g_bridge.scatter) rtnw::no_bridge("material::scatter");
    return rtnw::g_bridge.scatter(this, r_in, rec, attenuation, scattered);
}
vec3 material::emitted(float u, float v, const vec3& p) const {
    if (!rtnw::g_bridge.emitted) rtnw::no_bridge("material::emitted");
    return rtnw::g_bridge.emitted(this, u, v, p);
}
vec3 texture::value(float u, float v, const vec3& p) const {
    if (!rtnw::g_bridge.value) rtnw::no_bridge("texture::value");
    return rtnw::g_bridge.value(this, u, v, p);
}

// ---- perlin tables, PSC/perlin.h:82-111 -------------------------------------------------------------------------
vec3* perlin::ranvec = nullptr;
int* perlin::perm_x = nullptr;
int* perlin::perm_y = nullptr;
int* perlin::perm_z = nullptr;

static int* shuffled_identity() {
    int* p = new int[256];
    for (int i = 0; i < 256; ++i) p[i] = i;
    for (int i = 255; i > 0; --i) {  // Fisher-Yates with the reference's index draw
        const int target = int(drand48() * (i + 1));
        const int tmp = p[i];
        p[i] = p[target];
        p[target] = tmp;
    }
    return p;
}

void perlin::regenerate() {
    vec3* rv = new vec3[256];
    for (int i = 0; i < 256; ++i) {
        // g++ evaluates constructor arguments right to left; spelled out so the draw order does not depend on it
        // (matches the reference binary built with g++ on x86-64, the oracle's pinned toolchain).
        const double dz = drand48();
        const double dy = drand48();
        const double dx = drand48();
        rv[i] = unit_vector(vec3(-1 + 2 * dx, -1 + 2 * dy, -1 + 2 * dz));
    }
    ranvec = rv;
    perm_x = shuffled_identity();
    perm_y = shuffled_identity();
    perm_z = shuffled_identity();
}

void perlin::regenerate_readme() {
    vec3* rv = new vec3[256];
    for (int i = 0; i < 256; ++i) rv[i] = vec3(float(drand48()), 0, 0);  // float* p = new float[256]; p[i] = drand48()
    ranvec = rv;
    perm_x = shuffled_identity();
    perm_y = shuffled_identity();
    perm_z = shuffled_identity();
}

// ---- rotate_y, PSC/hitable.h:98-126 -----------------------------------------------------------------------------
rotate_y::rotate_y(hitable* p, float angle) : ptr(p) {
    const float radians = (M_PI / 180.) * angle;
    sin_theta = std::sin(radians);
    cos_theta = std::cos(radians);
    hasbox = ptr->bounding_box(0, 1, bbox);
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    float hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int corner = 0; corner < 8; ++corner) {  // i outermost, k innermost, as in the reference
        const int i = corner >> 2, j = (corner >> 1) & 1, k = corner & 1;
        const float x = i * bbox.max().x() + (1 - i) * bbox.min().x();
        const float y = j * bbox.max().y() + (1 - j) * bbox.min().y();
        const float z = k * bbox.max().z() + (1 - k) * bbox.min().z();
        const float rotated[3] = {cos_theta * x + sin_theta * z, y, -sin_theta * x + cos_theta * z};
        for (int c = 0; c < 3; ++c) {
            if (rotated[c] > hi[c]) hi[c] = rotated[c];
            if (rotated[c] < lo[c]) lo[c] = rotated[c];
        }
    }
    bbox = aabb(vec3(lo[0], lo[1], lo[2]), vec3(hi[0], hi[1], hi[2]));
}

// ---- hitable_list -----------------------------------------------------------------------------------------------
bool hitable_list::bounding_box(float t0, float t1, aabb& out) const {
    if (list_size < 1) return false;
    aabb acc;
    for (int i = 0; i < list_size; ++i) {
        aabb b;
        if (!list[i]->bounding_box(t0, t1, b)) return false;
        acc = (i == 0) ? b : surrounding_box(acc, b);
    }
    out = acc;
    return true;
}

// ---- box, PSC/box.h:23-34: +z, -z(flipped), +y, -y(flipped), +x, -x(flipped) -------------------------------------
box::box(const vec3& p0, const vec3& p1, material* m) : pmin(p0), pmax(p1), mat_ptr(m) {
    hitable** faces = new hitable*[6];
    faces[0] = new xy_rect(p0.x(), p1.x(), p0.y(), p1.y(), p1.z(), m);
    faces[1] = new flip_normals(new xy_rect(p0.x(), p1.x(), p0.y(), p1.y(), p0.z(), m));
    faces[2] = new xz_rect(p0.x(), p1.x(), p0.z(), p1.z(), p1.y(), m);
    faces[3] = new flip_normals(new xz_rect(p0.x(), p1.x(), p0.z(), p1.z(), p0.y(), m));
    faces[4] = new yz_rect(p0.y(), p1.y(), p0.z(), p1.z(), p1.x(), m);
    faces[5] = new flip_normals(new yz_rect(p0.y(), p1.y(), p0.z(), p1.z(), p0.x(), m));
    list_ptr = new hitable_list(faces, 6);
}

// ---- bvh_node, PSC/bvh.h:58-121 ---------------------------------------------------------------------------------
namespace {
template <int AXIS>
int cmp_box_min(const void* a, const void* b) {
    aabb ba, bb;
    const hitable* ha = *static_cast<hitable* const*>(a);
    const hitable* hb = *static_cast<hitable* const*>(b);
    if (!ha->bounding_box(0, 0, ba) || !hb->bounding_box(0, 0, bb)) std::cerr << "no bounding box in bvh_node constructor\n";
    // never returns 0, and compares the float difference against a double zero, as the reference does
    return (ba.min()[AXIS] - bb.min()[AXIS] < 0.0) ? -1 : 1;
}
}  // namespace

void bvh_node::build(hitable** l, int n) {
    const int axis = int(3 * drand48());
    qsort(l, n, sizeof(hitable*), axis == 0 ? cmp_box_min<0> : (axis == 1 ? cmp_box_min<1> : cmp_box_min<2>));
    if (n == 1) {
        left = right = l[0];
    } else if (n == 2) {
        left = l[0];
        right = l[1];
    } else {
        left = new bvh_node(l, n / 2, time0, time1, inner_tag());
        right = new bvh_node(l + n / 2, n - n / 2, time0, time1, inner_tag());
    }
    aabb bl, br;
    if (!left->bounding_box(time0, time1, bl) || !right->bounding_box(time0, time1, br))
        std::cerr << "no bounding box in bvh_node constructor\n";
    box = surrounding_box(bl, br);
}

bvh_node::bvh_node(hitable** l, int n, float t0, float t1) : left(nullptr), right(nullptr), time0(t0), time1(t1) {
    creation_order.assign(l, l + n);
    build(l, n);
}
bvh_node::bvh_node(hitable** l, int n, float t0, float t1, inner_tag) : left(nullptr), right(nullptr), time0(t0), time1(t1) {
    build(l, n);
}

ray camera::get_ray(float s, float t) const {
    if (!rtnw::g_bridge.get_ray) rtnw::no_bridge("camera::get_ray");
    return rtnw::g_bridge.get_ray(this, s, t);
}

// ---- camera, PSC/camera.h:21-39 ---------------------------------------------------------------------------------
camera::camera(vec3 lookfrom, vec3 lookat, vec3 vup, float vfov, float aspect, float aperture, float focus_dist, float t0, float t1) {
    time0 = t0;
    time1 = t1;
    len_radius = aperture / 2;
    const float theta = vfov * M_PI / 180;
    const float half_height = std::tan(theta / 2);
    const float half_width = aspect * half_height;
    origin = lookfrom;
    w = unit_vector(lookfrom - lookat);
    u = unit_vector(cross(vup, w));
    v = cross(w, u);
    lower_left_corner = origin - half_width * focus_dist * u - half_height * focus_dist * v - focus_dist * w;
    horizontal = 2 * half_width * focus_dist * u;
    vertical = 2 * half_height * focus_dist * v;
}
