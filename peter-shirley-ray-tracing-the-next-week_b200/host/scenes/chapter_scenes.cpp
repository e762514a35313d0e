#include "scenes/chapter_scenes.hpp"

// Draw order.  The reference writes e.g. `vec3(a+0.9*drand48(), 0.2, b+0.9*drand48())` (PSC/main.cpp:62), whose
// argument evaluation order is unspecified in C++; the reference binary the oracle pins (g++ 13, x86-64)
// evaluates call arguments right to left.  Every draw below is spelled out as a named statement in that order,
// so these builders produce the same scenes under any compiler (checked object-for-object against the reference
// build in tests/test_scene_parity.py).

namespace rtnw_scenes {
namespace {

struct pool {  // builders hand the list array to hitable_list, which keeps the pointer (as in the reference)
    hitable** items;
    int n;
    explicit pool(int capacity) : items(new hitable*[capacity]), n(0) {}
    void add(hitable* h) { items[n++] = h; }
    hitable* as_list() { return new hitable_list(items, n); }
};

material* matte(float r, float g, float b) { return new lambertian(new constant_texture(vec3(r, g, b))); }
material* lamp(float v) { return new diffuse_light(new constant_texture(vec3(v, v, v))); }

material* add_cornell_walls(pool& w, material* light, float lx0, float lx1, float lz0, float lz1) {
    material* red = matte(0.65, 0.05, 0.05);
    material* white = matte(0.73, 0.73, 0.73);
    material* green = matte(0.12, 0.45, 0.15);
    w.add(new flip_normals(new yz_rect(0, 555, 0, 555, 555, green)));
    w.add(new yz_rect(0, 555, 0, 555, 0, red));
    w.add(new xz_rect(lx0, lx1, lz0, lz1, 554, light));
    w.add(new flip_normals(new xz_rect(0, 555, 0, 555, 555, white)));
    w.add(new xz_rect(0, 555, 0, 555, 0, white));
    w.add(new flip_normals(new xy_rect(0, 555, 0, 555, 555, white)));
    return white;  // the blocks share the walls' white material, as in the reference
}

hitable* cornell_block(float height, float degrees, const vec3& where, material* m) {
    return new translate(new rotate_y(new box(vec3(0, 0, 0), vec3(165, height, 165), m), degrees), where);
}

}  // namespace

hitable* random_scene_ch01() {
    pool w(501);
    w.add(new sphere(vec3(0, -700, 0), 700, matte(0.5, 0.5, 0.5)));
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            const float choose_mat = drand48();
            const double dz = drand48();
            const double dx = drand48();
            const vec3 center(a + 0.9 * dx, 0.2, b + 0.9 * dz);
            if (!((center - vec3(4, 0.2, 0)).length() > 0.9)) continue;
            if (choose_mat < 0.8) {  // diffuse, moving upwards during the shutter interval
                const double b0 = drand48(), b1 = drand48();
                const double g0 = drand48(), g1 = drand48();
                const double r0 = drand48(), r1 = drand48();
                material* m = new lambertian(new constant_texture(vec3(r0 * r1, g0 * g1, b0 * b1)));
                const double lift = drand48();
                w.add(new moving_sphere(center, center + vec3(0, 0.5 * lift, 0), 0.0, 1.0, 0.2, m));
            } else if (choose_mat < 0.95) {  // metal
                const double fuzz = drand48();
                const double cb = drand48();
                const double cg = drand48();
                const double cr = drand48();
                w.add(new sphere(center, 0.2, new metal(vec3(0.5 * (1 + cr), 0.5 * (1 + cg), 0.5 * (1 + cb)), 0.5 * fuzz)));
            } else {  // glass
                w.add(new sphere(center, 0.2, new dielectric(1.5)));
            }
        }
    }
    w.add(new sphere(vec3(0, 1, 0), 1.0, new dielectric(2.5)));
    w.add(new sphere(vec3(-4, 1, 0), 1.0, matte(0.4, 0.2, 0.1)));
    w.add(new sphere(vec3(4, 1, 0), 1.0, new metal(vec3(1, 1, 1), 0.0)));
    return w.as_list();
}

// PSC/main.cpp:48-85, the live version: checker ground, the diffuse branch commented out (it draws nothing), metal and
// glass small spheres.  Draws spelled out in g++'s right-to-left argument order, as in random_scene_ch01().
hitable* random_scene() {
    pool w(501);
    texture* checker = new checker_texture(new constant_texture(vec3(0.2, 0.3, 0.1)), new constant_texture(vec3(0.9, 0.9, 0.9)));
    w.add(new sphere(vec3(0, -700, 0), 700, new lambertian(checker)));
    for (int a = -11; a < 11; a++) {
        for (int b = -11; b < 11; b++) {
            const float choose_mat = drand48();
            const double dz = drand48();
            const double dx = drand48();
            const vec3 center(a + 0.9 * dx, 0.2, b + 0.9 * dz);
            if (!((center - vec3(4, 0.2, 0)).length() > 0.9)) continue;
            if (choose_mat < 0.8) {
                // diffuse: commented out in the reference
            } else if (choose_mat < 0.95) {
                const double fuzz = drand48();
                const double cb = drand48();
                const double cg = drand48();
                const double cr = drand48();
                w.add(new sphere(center, 0.2, new metal(vec3(0.5 * (1 + cr), 0.5 * (1 + cg), 0.5 * (1 + cb)), 0.5 * fuzz)));
            } else {
                w.add(new sphere(center, 0.2, new dielectric(1.5)));
            }
        }
    }
    w.add(new sphere(vec3(0, 1, 0), 1.0, new dielectric(2.5)));
    w.add(new sphere(vec3(-4, 1, 0), 1.0, matte(0.4, 0.2, 0.1)));
    w.add(new sphere(vec3(4, 1, 0), 1.0, new metal(vec3(1, 1, 1), 0.0)));
    return w.as_list();
}

// PSC/main.cpp:135-145
hitable* test_scene() {
    texture* checker = new checker_texture(new constant_texture(vec3(0.2, 0.3, 0.1)), new constant_texture(vec3(0.9, 0.9, 0.9)));
    pool w(3);
    w.add(new sphere(vec3(0, -700, 0), 700, new lambertian(checker)));
    w.add(new sphere(vec3(0, 2, 0), 2, new lambertian(new noise_texture(4))));
    w.add(new sphere(vec3(0, 7, 0), 2, lamp(11)));
    return w.as_list();
}

hitable* two_perlin_spheres() {
    texture* checker = new checker_texture(new constant_texture(vec3(0.2, 0.3, 0.1)), new constant_texture(vec3(0.9, 0.9, 0.9)));
    pool w(2);
    w.add(new sphere(vec3(0, -1000, 0), 1000, new lambertian(checker)));
    w.add(new sphere(vec3(0, 2, 0), 2, new lambertian(new noise_texture(4))));
    return w.as_list();
}

hitable* cornell_box() {
    pool w(8);
    material* white = add_cornell_walls(w, lamp(15), 213, 343, 227, 332);
    w.add(cornell_block(165, -18, vec3(130, 0, 65), white));
    w.add(cornell_block(330, 15, vec3(265, 0, 295), white));
    return w.as_list();
}

hitable* cornell_smoke() {
    pool w(8);
    material* white = add_cornell_walls(w, lamp(4), 113, 443, 127, 432);
    hitable* b1 = cornell_block(165, -18, vec3(130, 0, 65), white);
    hitable* b2 = cornell_block(330, 15, vec3(265, 0, 295), white);
    w.add(new constant_medium(b1, 0.01, new constant_texture(vec3(1.0, 1.0, 1.0))));
    w.add(new constant_medium(b2, 0.01, new constant_texture(vec3(0.0, 0.0, 0.0))));
    return w.as_list();
}

namespace {
// the floor of boxes with random heights, PSC/main.cpp:196-208 (one draw per box)
void add_floor(pool& dst, int nb, float w, material* ground) {
    for (int i = 0; i < nb; i++) {
        for (int j = 0; j < nb; j++) {
            const float x0 = i * w, z0 = j * w, y0 = 0;
            const float x1 = x0 + w;
            const float y1 = 100 * (drand48() + 0.01);
            const float z1 = z0 + w;
            dst.add(new box(vec3(x0, y0, z0), vec3(x1, y1, z1), ground));
        }
    }
}

// light, moving sphere, glass, metal, two media: PSC/main.cpp:210-222
void add_final_props(pool& w) {
    w.add(new xz_rect(123, 423, 147, 412, 554, lamp(7)));
    const vec3 center(400, 400, 200);
    w.add(new moving_sphere(center, center + vec3(30, 0, 0), 0, 1, 50, matte(0.7, 0.3, 0.1)));
    w.add(new sphere(vec3(260, 150, 45), 50, new dielectric(1.5)));
    w.add(new sphere(vec3(0, 150, 145), 50, new metal(vec3(0.8, 0.8, 0.9), 10.0)));
    hitable* boundary = new sphere(vec3(360, 150, 145), 70, new dielectric(1.5));
    w.add(boundary);
    w.add(new constant_medium(boundary, 0.2, new constant_texture(vec3(0.2, 0.4, 0.9))));
    boundary = new sphere(vec3(0, 0, 0), 5000, new dielectric(1.5));
    w.add(new constant_medium(boundary, 0.0001, new constant_texture(vec3(1.0, 1.0, 1.0))));
}
}  // namespace

hitable* final_scene() {
    pool w(3000);
    material* white = matte(0.73, 0.73, 0.73);
    material* ground = matte(0.48, 0.83, 0.53);
    add_floor(w, 10, 100, ground);
    add_final_props(w);
    w.add(new sphere(vec3(220, 280, 300), 80, new lambertian(new noise_texture(0.1))));
    for (int j = 0; j < 1000; j++) {
        const double dz = drand48(), dy = drand48(), dx = drand48();
        w.add(new sphere(vec3(165 * dx - 100, 165 * dy + 270, 165 * dz + 395), 10, white));
    }
    return w.as_list();
}

unsigned char* synthetic_earth(int& nx, int& ny) {
    nx = 1024;
    ny = 512;
    unsigned char* px = new unsigned char[(size_t)nx * ny * 3];
    for (int y = 0; y < ny; ++y) {
        for (int x = 0; x < nx; ++x) {
            // integer-only pattern (blocky "continents" over a banded "ocean") so every toolchain produces the same bytes
            const unsigned h = (unsigned)(x / 32) * 2654435761u ^ (unsigned)(y / 32) * 40503u;
            const bool land = ((h >> 13) & 7u) < 3u;
            unsigned char* p = px + 3 * ((size_t)y * nx + x);
            p[0] = (unsigned char)(land ? 60 + ((x * 5 + y * 3) & 63) : 10 + (y & 31));
            p[1] = (unsigned char)(land ? 120 + ((x * 3 + y * 7) & 63) : 40 + ((x + y) & 63));
            p[2] = (unsigned char)(land ? 40 + ((x ^ y) & 31) : 140 + ((x * 2 + y) & 63));
        }
    }
    return px;
}

hitable* earth(unsigned char* tex, int nx, int ny) {
    pool w(2);
    w.add(new xz_rect(63, 483, 55, 482, 554, lamp(7)));
    w.add(new sphere(vec3(360, 250, 150), 100, new lambertian(new image_texture(tex, nx, ny))));
    return w.as_list();
}
hitable* earth() {
    int nx, ny;
    unsigned char* tex = synthetic_earth(nx, ny);
    return earth(tex, nx, ny);
}

hitable* final_northstar() {
    pool w(30);
    material* white = matte(0.73, 0.73, 0.73);
    material* ground = matte(0.48, 0.83, 0.53);
    const int nb = 32;
    pool floor(nb * nb);
    add_floor(floor, nb, 1000.0f / nb, ground);
    w.add(new bvh_node(floor.items, floor.n, 0, 1));
    add_final_props(w);
    int nx, ny;
    unsigned char* tex = synthetic_earth(nx, ny);
    w.add(new sphere(vec3(400, 200, 400), 100, new lambertian(new image_texture(tex, nx, ny))));
    w.add(new sphere(vec3(220, 280, 300), 80, new lambertian(new noise_texture(0.1))));
    pool cluster(1000);
    for (int j = 0; j < 1000; j++) {
        const double dz = drand48(), dy = drand48(), dx = drand48();
        cluster.add(new sphere(vec3(165 * dx, 165 * dy, 165 * dz), 10, white));
    }
    w.add(new translate(new rotate_y(new bvh_node(cluster.items, cluster.n, 0.0, 1.0), 15), vec3(-100, 270, 395)));
    return w.as_list();
}

hitable* simple_light() {
    texture* checker = new checker_texture(new constant_texture(vec3(0.2, 0.3, 0.1)), new constant_texture(vec3(0.9, 0.9, 0.9)));
    pool w(4);
    w.add(new sphere(vec3(0, 2, 0), 2, new lambertian(new noise_texture(4))));
    w.add(new sphere(vec3(0, -700, 0), 700, new lambertian(checker)));
    w.add(new sphere(vec3(0, 7, 0), 2, lamp(4)));
    w.add(new xy_rect(3, 5, 1, 3, -2, lamp(4)));
    return w.as_list();
}

hitable* two_spheres() {
    material* red = matte(0.65, 0.05, 0.05);
    pool w(2);
    w.add(new sphere(vec3(0, -10, 0), 10, red));
    w.add(new yz_rect(0, 555, 0, 555, 0, red));
    return w.as_list();
}

// Not a reference scene: a queue stress fixture for the tests.  600 concentric glass/diffuse shells, so a ray aimed at the
// centre passes every bounding box of any bvh_node built over them (hundreds of leaves per ray, all with nearly equal t).
hitable* stress_shells() {
    pool w(600);
    for (int k = 0; k < 600; ++k) {
        material* m = (k % 3 == 0) ? static_cast<material*>(new dielectric(1.5f)) : matte(0.2f + 0.001f * k, 0.5f, 0.9f - 0.001f * k);
        w.add(new sphere(vec3(0.01f * (k % 7), 0.02f * (k % 5), 0.0f), 1.0f + 0.01f * k, m));
    }
    return w.as_list();
}

// Not a reference scene: two bvh_nodes that are NEIGHBOURS in the top-level list (and a third after one sphere), so the
// cooperative traversal enters a BVH item straight after leaving one (tests/test_gpu_parity.py).
hitable* twin_bvh() {
    pool w(4);
    for (int part = 0; part < 3; ++part) {
        pool g(40);
        for (int k = 0; k < 40; ++k) {
            const float x = -6.f + 0.3f * k + 0.05f * part, z = -3.f + 2.5f * part + 0.4f * (k % 5), r = 0.15f + 0.01f * (k % 7);
            material* m = (k % 4 == 0) ? static_cast<material*>(new metal(vec3(0.8f, 0.7f, 0.6f), 0.1f)) : matte(0.3f + 0.01f * k, 0.4f, 0.2f + 0.2f * part);
            g.add(new sphere(vec3(x, 0.2f + 0.1f * (k % 3), z), r, m));
        }
        if (part == 2) w.add(new sphere(vec3(0, -700, 0), 700, matte(0.5, 0.5, 0.5)));
        w.add(new bvh_node(g.items, g.n, 0, 1));
    }
    return w.as_list();
}

hitable* wrap_in_bvh(hitable* flat_list, float t0, float t1) {
    hitable_list* l = static_cast<hitable_list*>(flat_list);
    return new bvh_node(l->list, l->list_size, t0, t1);
}

view view_ch01() { return view{vec3(13, 2, 3), vec3(0, 0, 0), 20, 0.1f, 10, 0, 1, 200, 100, 100, 0.001f, true, false, false}; }
view view_two_perlin() { return view{vec3(13, 2, 3), vec3(0, 0, 0), 20, 0.0f, 10, 0, 1, 400, 200, 256, 0.001f, true, false, false}; }
view view_cornell() { return view{vec3(278, 278, -800), vec3(278, 278, 0), 40, 0.0f, 10, 0, 1, 500, 500, 1000, 0.001f, false, true, true}; }
view view_final() { return view{vec3(228, 278, -800), vec3(278, 278, 0), 40, 0.0f, 10, 0, 1, 1000, 1000, 100, 0.001f, false, true, true}; }

camera make_camera(const view& v, int nx, int ny) {
    return camera(v.lookfrom, v.lookat, vec3(0, 1, 0), v.vfov, float(nx) / float(ny), v.aperture, v.focus_dist, v.time0, v.time1);
}

}  // namespace rtnw_scenes
