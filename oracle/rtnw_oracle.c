/*
 * oracle/rtnw_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C, single-threaded restatement of the reference's per-pixel path-tracing sample loop
 * (EStormLynn/Peter-Shirley-Ray-Tracing-the-next-week; PSC/ = "Peter-Shirley-Project Code/"), written against the
 * flattened tables of include/rtnw.h so that it checks the flattener and the CUDA path on exactly the data the GPU
 * sees.  It follows the reference's own control structure — recursive bvh_node::hit, narrowing hitable_list::hit,
 * recursive color() — and its float/double promotions expression by expression; each function cites the lines it
 * restates.  With the same toolchain (gcc, -O2 -ffp-contract=off, glibc libm) it is bit-identical to the reference,
 * which tests/test_oracle_pinning.py asserts against oracle/_ref/libref_oracle.so (the reference itself, compiled).
 *
 * PARITY PINNED: yes — against the reference renderer run in this container (closest hits, scatter, textures, perlin,
 * camera rays and whole images under the shared sample stream) and against the committed fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * Deviations from the reference shared with the product: aabb::hit uses r.origin() (F2);
 * moving spheres and media report u = v = 0 (the reference leaves them unwritten, F5); the sample stream is the
 * framework's (Philox4x32-10 seeding the drand48 recurrence per path; media draw keyed Philox numbers).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "rtnw.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { float e[3]; } vec3;
typedef struct { vec3 A, B; float time; } ray;
typedef struct { float t, u, v; vec3 p, normal; int mat; int leaf; int face; } hit_record;

/* ---- PSC/vec3.h:61-145 ------------------------------------------------------------------------------------ */
static vec3 V(float a, float b, float c) { vec3 r; r.e[0] = a; r.e[1] = b; r.e[2] = c; return r; }
static vec3 vadd(vec3 a, vec3 b) { return V(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
static vec3 vsub(vec3 a, vec3 b) { return V(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
static vec3 vmul(vec3 a, vec3 b) { return V(a.e[0] * b.e[0], a.e[1] * b.e[1], a.e[2] * b.e[2]); }
static vec3 smul(float t, vec3 a) { return V(t * a.e[0], t * a.e[1], t * a.e[2]); }
static vec3 sdiv(vec3 a, float t) { return V(a.e[0] / t, a.e[1] / t, a.e[2] / t); }
static vec3 vneg(vec3 a) { return V(-a.e[0], -a.e[1], -a.e[2]); }
static float dot(vec3 a, vec3 b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }
static float length(vec3 a) { return sqrtf(a.e[0] * a.e[0] + a.e[1] * a.e[1] + a.e[2] * a.e[2]); }
static vec3 unit_vector(vec3 a) { return sdiv(a, length(a)); }
static vec3 point_at(const ray* r, float t) { return vadd(r->A, smul(t, r->B)); } /* PSC/ray.h:18 */

/* ---- sample stream, shared with oracle/ref_harness.cpp and csrc/rtnw_device.cuh ------------- */
typedef struct {
    uint32_t key[2], pixel, sample;
    uint64_t x;
    int seeded;
    int depth;          /* index of the current top-level closest-hit query of the path */
    int in_boundary;    /* > 0 while a constant_medium probes its boundary (those hit() calls are not counted as tests) */
    uint64_t rays, box_tests, prim_tests;
} stream;

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u01(uint32_t x) { return (double)((float)(x >> 8) * (1.0f / 16777216.0f)); }

static void begin_path(stream* g, uint32_t pixel, uint32_t sample) {
    g->pixel = pixel; g->sample = sample; g->seeded = 0; g->depth = -1;
}
/* stands where the reference calls drand48() outside constant_medium::hit */
static double draw(stream* g) {
    if (!g->seeded) {
        const uint32_t ctr[4] = {0u, 0u, g->sample, g->pixel};
        uint32_t o[4];
        philox4x32_10(ctr, g->key, o);
        g->x = ((uint64_t)(o[1] & 0xffffu) << 32) | (uint64_t)o[0];
        g->seeded = 1;
    }
    g->x = (g->x * 0x5DEECE66DULL + 0xBULL) & 0xffffffffffffULL;
    return (double)((float)(uint32_t)(g->x >> 24) * (1.0f / 16777216.0f));
}
/* stands where constant_medium::hit calls drand48(): keyed by the medium's leaf id and the ray's depth in the path */
static double keyed_draw(const stream* g, int leaf) {
    const uint32_t ctr[4] = {(uint32_t)leaf, 1u + (uint32_t)g->depth, g->sample, g->pixel};
    uint32_t o[4];
    philox4x32_10(ctr, g->key, o);
    return u01(o[0]);
}

/* PSC/material.h:41-47; g++ evaluates vec3(drand48(),drand48(),drand48()) right to left: first draw -> z */
static vec3 random_in_unit_sphere(stream* g) {
    vec3 p;
    do {
        const double dz = draw(g), dy = draw(g), dx = draw(g);
        p = vsub(smul(2.0, V(dx, dy, dz)), V(1, 1, 1));
    } while (dot(p, p) >= 1.0);
    return p;
}

typedef struct { const rtnw_scene_desc* d; stream* g; } ctx;

/* ---- wrappers: PSC/hitable.h:39-150 ----------------------------------------------------------------------- */
static int chain_len(const rtnw_scene_desc* d, uint32_t chain) { return chain ? (int)(d->xforms[chain].kind >> 8) : 0; }

static ray xform_ray(const rtnw_scene_desc* d, uint32_t chain, ray r) {
    const int n = chain_len(d, chain);
    for (int k = 0; k < n; ++k) {
        const rtnw_xform_op* op = &d->xforms[chain + k];
        if ((op->kind & 0xffu) == RTNW_XF_TRANSLATE) {      /* translate::hit, :66-68 */
            r.A = vsub(r.A, V(op->a, op->b, op->c));
        } else {                                            /* rotate_y::hit, :128-135; a = sin, b = cos */
            vec3 origin = r.A, direction = r.B;
            origin.e[0] = op->b * r.A.e[0] - op->a * r.A.e[2];
            origin.e[2] = op->a * r.A.e[0] + op->b * r.A.e[2];
            direction.e[0] = op->b * r.B.e[0] - op->a * r.B.e[2];
            direction.e[2] = op->a * r.B.e[0] + op->b * r.B.e[2];
            r.A = origin; r.B = direction;
        }
    }
    return r;
}
static void xform_back(const rtnw_scene_desc* d, uint32_t chain, hit_record* rec) {
    for (int k = chain_len(d, chain) - 1; k >= 0; --k) {
        const rtnw_xform_op* op = &d->xforms[chain + k];
        if ((op->kind & 0xffu) == RTNW_XF_TRANSLATE) {      /* :69-70 */
            rec->p = vadd(rec->p, V(op->a, op->b, op->c));
        } else {                                            /* :137-146 */
            vec3 p = rec->p, normal = rec->normal;
            p.e[0] = op->b * rec->p.e[0] + op->a * rec->p.e[2];
            p.e[2] = -op->a * rec->p.e[0] + op->b * rec->p.e[2];
            normal.e[0] = op->b * rec->normal.e[0] + op->a * rec->normal.e[2];
            normal.e[2] = -op->a * rec->normal.e[0] + op->b * rec->normal.e[2];
            rec->p = p; rec->normal = normal;
        }
    }
}

/* ---- primitives ------------------------------------------------------------------------------------------- */
static void get_sphere_uv(vec3 p, float* u, float* v) { /* PSC/hitable.h:14-19 */
    float phi = atan2f(p.e[2], p.e[0]);
    float theta = asinf(p.e[1]);
    *u = 1 - (phi + M_PI) / (2 * M_PI);
    *v = (theta + M_PI / 2) / M_PI;
}
/* PSC/sphere.h:25-52 (with_uv) and :92-118 (moving, center already evaluated at r.time()) */
static int sphere_hit(vec3 center, float radius, int with_uv, const ray* r, float t_min, float t_max, hit_record* rec) {
    vec3 oc = vsub(r->A, center);
    float a = dot(r->B, r->B);
    float b = dot(oc, r->B);
    float c = dot(oc, oc) - radius * radius;
    float discriminant = b * b - a * c;
    if (discriminant > 0) {
        float temp = (-b - sqrtf(discriminant)) / a;
        for (int root = 0; root < 2; ++root) {
            if (temp < t_max && temp > t_min) {
                rec->t = temp;
                rec->p = point_at(r, rec->t);
                rec->normal = sdiv(vsub(rec->p, center), radius);
                rec->u = 0; rec->v = 0;
                if (with_uv) get_sphere_uv(sdiv(vsub(rec->p, center), radius), &rec->u, &rec->v);
                return 1;
            }
            temp = (-b + sqrtf(discriminant)) / a;
        }
    }
    return 0;
}
/* PSC/aarect.h:50-100; axis n = plane normal, (a, b) = extent axes */
static int rect_hit(int n, int a, int b, float a0, float a1, float b0, float b1, float k, const ray* r, float t0, float t1,
                    hit_record* rec) {
    float t = (k - r->A.e[n]) / r->B.e[n];
    if (t < t0 || t > t1) return 0;
    float x = r->A.e[a] + t * r->B.e[a];
    float y = r->A.e[b] + t * r->B.e[b];
    if (x < a0 || x > a1 || y < b0 || y > b1) return 0;
    rec->u = (x - a0) / (a1 - a0);
    rec->v = (y - b0) / (b1 - b0);
    rec->t = t;
    rec->p = point_at(r, t);
    rec->normal = V(n == 0, n == 1, n == 2);
    return 1;
}
/* PSC/box.h:23-38: inner hitable_list of +z, -z(flipped), +y, -y(flipped), +x, -x(flipped) */
static int box_hit(const float* f, const ray* r, float t0, float t1, hit_record* rec) {
    hit_record temp;
    int hit_anything = 0;
    double closest_so_far = t1;
    for (int face = 0; face < 6; ++face) {
        int h;
        const float lo = f[face >> 1 == 0 ? 2 : (face >> 1 == 1 ? 1 : 0)], hi = f[3 + (face >> 1 == 0 ? 2 : (face >> 1 == 1 ? 1 : 0))];
        const float k = (face & 1) ? lo : hi;
        if (face < 2) h = rect_hit(2, 0, 1, f[0], f[3], f[1], f[4], k, r, t0, closest_so_far, &temp);
        else if (face < 4) h = rect_hit(1, 0, 2, f[0], f[3], f[2], f[5], k, r, t0, closest_so_far, &temp);
        else h = rect_hit(0, 1, 2, f[1], f[4], f[2], f[5], k, r, t0, closest_so_far, &temp);
        if (h) {
            if (face & 1) temp.normal = vneg(temp.normal); /* flip_normals, PSC/hitable.h:42-49 */
            hit_anything = 1;
            closest_so_far = temp.t;
            temp.face = face;
            *rec = temp;
        }
    }
    return hit_anything;
}

static int slots_hit(const ctx* c, int first, int count, const ray* r, float t_min, float t_max, hit_record* rec);

/* one prim slot = one hitable handed to a list/BVH, with its wrappers; *used = slots it occupies */
static int prim_hit(const ctx* c, int s, const ray* r_outer, float t_min, float t_max, hit_record* rec, int* used) {
    const rtnw_scene_desc* d = c->d;
    const rtnw_prim* p = &d->prims[s];
    const uint32_t kind = RTNW_KX_KIND(p->kx), chain = RTNW_KX_XFORM(p->kx);
    const ray r = xform_ray(d, chain, *r_outer);
    int h = 0;
    *used = 1;
    rec->face = 0;
    if (!c->g->in_boundary) c->g->prim_tests++;
    switch (kind) {
        case RTNW_PRIM_SPHERE: h = sphere_hit(V(p->f[0], p->f[1], p->f[2]), p->f[3], 1, &r, t_min, t_max, rec); break;
        case RTNW_PRIM_MOVING_SPHERE: { /* PSC/sphere.h:81-83 */
            const rtnw_prim* e = &d->prims[s + 1];
            vec3 c0 = V(p->f[0], p->f[1], p->f[2]), c1 = V(e->f[0], e->f[1], e->f[2]);
            vec3 center = vadd(c0, smul((r.time - p->f[4]) / (p->f[5] - p->f[4]), vsub(c1, c0)));
            h = sphere_hit(center, p->f[3], 0, &r, t_min, t_max, rec);
            *used = 2;
            break;
        }
        case RTNW_PRIM_RECT_XY: h = rect_hit(2, 0, 1, p->f[0], p->f[1], p->f[2], p->f[3], p->f[4], &r, t_min, t_max, rec); break;
        case RTNW_PRIM_RECT_XZ: h = rect_hit(1, 0, 2, p->f[0], p->f[1], p->f[2], p->f[3], p->f[4], &r, t_min, t_max, rec); break;
        case RTNW_PRIM_RECT_YZ: h = rect_hit(0, 1, 2, p->f[0], p->f[1], p->f[2], p->f[3], p->f[4], &r, t_min, t_max, rec); break;
        case RTNW_PRIM_BOX: h = box_hit(p->f, &r, t_min, t_max, rec); break;
        case RTNW_PRIM_MEDIUM: { /* PSC/constant_medium.h:26-50 */
            int bfirst, bcount, leaf;
            memcpy(&bfirst, &p->f[1], 4); memcpy(&bcount, &p->f[2], 4); memcpy(&leaf, &p->f[3], 4);
            hit_record rec1, rec2;
            c->g->in_boundary++;
            const int h1 = slots_hit(c, bfirst, bcount, &r, -FLT_MAX, FLT_MAX, &rec1);
            const int h2 = h1 && slots_hit(c, bfirst, bcount, &r, rec1.t + 0.0001, FLT_MAX, &rec2);
            c->g->in_boundary--;
            if (h1) {
                if (h2) {
                    if (rec1.t < t_min) rec1.t = t_min;
                    if (rec2.t > t_max) rec2.t = t_max;
                    if (rec1.t >= rec2.t) break;
                    if (rec1.t < 0) rec1.t = 0;
                    float distance_inside_boundary = (rec2.t - rec1.t) * length(r.B);
                    float hit_distance = -(1 / p->f[0]) * log(keyed_draw(c->g, leaf));
                    if (hit_distance < distance_inside_boundary) {
                        rec->t = rec1.t + hit_distance / length(r.B);
                        rec->p = point_at(&r, rec->t);
                        rec->normal = V(1, 0, 0);
                        rec->u = 0; rec->v = 0;
                        h = 1;
                    }
                }
            }
            break;
        }
        default: break;
    }
    if (h) {
        if (RTNW_KX_FLIP(p->kx)) rec->normal = vneg(rec->normal);
        xform_back(d, chain, rec);
        rec->mat = p->mat;
        rec->leaf = s;
    }
    return h;
}

/* hitable_list::hit over a slot range, PSC/hitable_list.h:20-32 */
static int slots_hit(const ctx* c, int first, int count, const ray* r, float t_min, float t_max, hit_record* rec) {
    hit_record temp_rec;
    int hit_anything = 0;
    double closest_so_far = t_max;
    for (int s = first; s < first + count;) {
        int used;
        if (prim_hit(c, s, r, t_min, closest_so_far, &temp_rec, &used)) {
            hit_anything = 1;
            closest_so_far = temp_rec.t;
            *rec = temp_rec;
        }
        s += used;
    }
    return hit_anything;
}

/* aabb::hit, PSC/aabb.h:33-49, with r.origin() in the two subtractions (F2) */
static int aabb_hit(const float* bmin, const float* bmax, const ray* r, float tmin, float tmax, stream* g) {
    g->box_tests++;
    for (int a = 0; a < 3; a++) {
        float invD = 1.0f / r->B.e[a];
        float t0 = (bmin[a] - r->A.e[a]) * invD;
        float t1 = (bmax[a] - r->A.e[a]) * invD;
        if (invD < 0.0f) { float tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin;
        tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return 0;
    }
    return 1;
}

/* bvh_node::hit, PSC/bvh.h:29-54, for node `idx` whose own box is [bmin,bmax] */
static int bvh_hit(const ctx* c, int idx, const float* bmin, const float* bmax, const ray* r, float tmin, float tmax, hit_record* rec) {
    if (!aabb_hit(bmin, bmax, r, tmin, tmax, c->g)) return 0;
    const rtnw_bvh_node* n = &c->d->nodes[idx];
    hit_record left_rec, right_rec;
    int hit_left, hit_right;
    if (n->left >= 0) hit_left = bvh_hit(c, n->left, n->lmin, n->lmax, r, tmin, tmax, &left_rec);
    else hit_left = slots_hit(c, ~n->left, n->lcount, r, tmin, tmax, &left_rec);
    if (n->right == RTNW_REF_NONE) { /* n == 1: right == left (PSC/bvh.h:106-108); the second call returns the same record */
        hit_right = hit_left;
        right_rec = left_rec;
    } else if (n->right >= 0) {
        hit_right = bvh_hit(c, n->right, n->rmin, n->rmax, r, tmin, tmax, &right_rec);
    } else {
        hit_right = slots_hit(c, ~n->right, n->rcount, r, tmin, tmax, &right_rec);
    }
    if (hit_left && hit_right) {
        if (left_rec.t < right_rec.t) *rec = left_rec;
        else *rec = right_rec;
        return 1;
    } else if (hit_left) {
        *rec = left_rec;
        return 1;
    } else if (hit_right) {
        *rec = right_rec;
        return 1;
    }
    return 0;
}

/* world->hit(r, t_min, t_max, rec), PSC/main.cpp:27: the top-level hitable_list over the items */
static int world_hit(const ctx* c, const ray* r, float t_min, float t_max, hit_record* rec) {
    const rtnw_scene_desc* d = c->d;
    hit_record temp_rec;
    int hit_anything = 0;
    double closest_so_far = t_max;
    c->g->rays++;
    c->g->depth++;
    for (int i = 0; i < d->n_items; ++i) {
        const rtnw_item* it = &d->items[i];
        const ray ri = xform_ray(d, it->xform, *r);
        int h;
        if (it->kind == RTNW_ITEM_BVH) h = bvh_hit(c, it->first, it->bmin, it->bmax, &ri, t_min, closest_so_far, &temp_rec);
        else h = slots_hit(c, it->first, it->count, &ri, t_min, closest_so_far, &temp_rec);
        if (h) {
            xform_back(d, it->xform, &temp_rec);
            hit_anything = 1;
            closest_so_far = temp_rec.t;
            *rec = temp_rec;
        }
    }
    return hit_anything;
}

/* ---- textures: PSC/perlin.h:25-74, PSC/texture.h:22-56, PSC/surface_texture.h:19-30 --------------------- */
static float perlin_interp(vec3 c[2][2][2], float u, float v, float w) {
    float uu = u * u * (3 - 2 * u);
    float vv = v * v * (3 - 2 * v);
    float ww = w * w * (3 - 2 * w);
    float accum = 0;
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) {
                vec3 weight_v = V(u - i, v - j, w - k);
                accum += (i * uu + (1 - i) * (1 - uu)) * (j * vv + (1 - j) * (1 - vv)) * (k * ww + (1 - k) * (1 - ww)) *
                         dot(c[i][j][k], weight_v);
            }
    return accum;
}
static float perlin_noise(const rtnw_scene_desc* d, vec3 p) {
    float u = p.e[0] - floorf(p.e[0]);
    float v = p.e[1] - floorf(p.e[1]);
    float w = p.e[2] - floorf(p.e[2]);
    u = u * u * (3 - 2 * u);
    v = v * v * (3 - 2 * v);
    w = w * w * (3 - 2 * w);
    int i = floorf(p.e[0]);
    int j = floorf(p.e[1]);
    int k = floorf(p.e[2]);
    vec3 c[2][2][2];
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                const int idx = d->perlin_perm_x[(i + di) & 255] ^ d->perlin_perm_y[(j + dj) & 255] ^ d->perlin_perm_z[(k + dk) & 255];
                c[di][dj][dk] = V(d->perlin_ranvec[3 * idx], d->perlin_ranvec[3 * idx + 1], d->perlin_ranvec[3 * idx + 2]);
            }
    return perlin_interp(c, u, v, w);
}
static float perlin_turb(const rtnw_scene_desc* d, vec3 p) {
    float accum = 0;
    vec3 temp_p = p;
    float weight = 1.0;
    for (int i = 0; i < 7; i++) {
        accum += weight * perlin_noise(d, temp_p);
        weight *= 0.5;
        temp_p = smul(2, temp_p);
    }
    return fabsf(accum);
}
/* README.md:516-630: the Chapter 4 noise functions before perlin.h's shipped form; ranfloat[i] = perlin_ranvec[3*i] */
static float readme_noise(const rtnw_scene_desc* d, uint32_t kind, vec3 p) {
    const float* rf = d->perlin_ranvec;
    if (kind == RTNW_TEX_NOISE_HASH) {
        int i = (int)(4 * p.e[0]) & 255;
        int j = (int)(4 * p.e[1]) & 255;
        int k = (int)(4 * p.e[2]) & 255;
        return rf[3 * (d->perlin_perm_x[i] ^ d->perlin_perm_y[j] ^ d->perlin_perm_z[k])];
    }
    float u = p.e[0] - floorf(p.e[0]);
    float v = p.e[1] - floorf(p.e[1]);
    float w = p.e[2] - floorf(p.e[2]);
    if (kind == RTNW_TEX_NOISE_HERMITE) {
        u = u * u * (3 - 2 * u);
        v = v * v * (3 - 2 * v);
        w = w * w * (3 - 2 * w);
    }
    int i = floorf(p.e[0]);
    int j = floorf(p.e[1]);
    int k = floorf(p.e[2]);
    float accum = 0;
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                const float c = rf[3 * (d->perlin_perm_x[(i + di) & 255] ^ d->perlin_perm_y[(j + dj) & 255] ^ d->perlin_perm_z[(k + dk) & 255])];
                accum += (di * u + (1 - di) * (1 - u)) * (dj * v + (1 - dj) * (1 - v)) * (dk * w + (1 - dk) * (1 - w)) * c;
            }
    return accum;
}

static vec3 texture_value(const rtnw_scene_desc* d, int tex, float u, float v, vec3 p) {
    const rtnw_texture* t = &d->textures[tex];
    switch (t->kind) {
        case RTNW_TEX_CONSTANT: return V(t->c[0], t->c[1], t->c[2]);
        case RTNW_TEX_CHECKER: {
            float sines = sinf(10 * p.e[0]) * sinf(10 * p.e[1]) * sinf(10 * p.e[2]);
            if (sines < 0) return texture_value(d, t->i1 /* odd */, u, v, p);
            return texture_value(d, t->i0 /* even */, u, v, p);
        }
        case RTNW_TEX_NOISE_HASH:
        case RTNW_TEX_NOISE_TRILINEAR:
        case RTNW_TEX_NOISE_HERMITE: {
            const float n = readme_noise(d, t->kind, p);
            return smul(n, V(1, 1, 1));
        }
        case RTNW_TEX_NOISE: {
            const float scale = t->c[0];
            return smul(1 + sinf(scale * p.e[0] + 5 * perlin_turb(d, smul(scale, p))), smul(0.5, V(1, 1, 1)));
        }
        default: {
            const int nx = t->i1, ny = t->i2;
            const uint8_t* data = d->images + t->i0;
            if (t->flags & RTNW_TEXF_BILINEAR) { /* this framework's option, not in the reference: same float32 steps as the device */
                const float fx = (1.f - u) * (float)nx - 0.5f, fy = (1.f - v) * (float)ny - 0.5f;
                const float x0 = floorf(fx), y0 = floorf(fy);
                const float wx = fx - x0, wy = fy - y0;
                int ii[4], jj[4];
                ii[0] = ii[2] = (int)x0; ii[1] = ii[3] = (int)x0 + 1;
                jj[0] = jj[1] = (int)y0; jj[2] = jj[3] = (int)y0 + 1;
                vec3 c[4];
                for (int q = 0; q < 4; ++q) {
                    const int a = ii[q] < 0 ? 0 : (ii[q] > nx - 1 ? nx - 1 : ii[q]), b = jj[q] < 0 ? 0 : (jj[q] > ny - 1 ? ny - 1 : jj[q]);
                    const uint8_t* px = data + 3 * a + 3 * nx * b;
                    c[q] = V((float)px[0] / 255.0f, (float)px[1] / 255.0f, (float)px[2] / 255.0f);
                }
                const vec3 top = vadd(smul(1.f - wx, c[0]), smul(wx, c[1])), bot = vadd(smul(1.f - wx, c[2]), smul(wx, c[3]));
                return vadd(smul(1.f - wy, top), smul(wy, bot));
            }
            int i = (1 - u) * nx;
            int j = (1 - v) * ny - 0.001;
            if (i < 0) i = 0;
            if (j < 0) j = 0;
            if (i > nx - 1) i = nx - 1;
            if (j > ny - 1) j = ny - 1;
            float r = (int)(data[3 * i + 3 * nx * j]) / 255.0;
            float g = (int)(data[3 * i + 3 * nx * j + 1]) / 255.0;
            float b = (int)(data[3 * i + 3 * nx * j + 2]) / 255.0;
            return V(r, g, b);
        }
    }
}

/* ---- materials: PSC/material.h:16-151 ------------------------------------------------------------------- */
static float schlick(float cosine, float ref_idx) {
    float r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * pow((1 - cosine), 5);
}
static int refract(vec3 v, vec3 n, float ni_over_nt, vec3* refracted) {
    vec3 uv = unit_vector(v);
    float dt = dot(uv, n);
    float discriminant = 1.0 - ni_over_nt * ni_over_nt * (1 - dt * dt);
    if (discriminant > 0) {
        *refracted = vsub(smul(ni_over_nt, vsub(uv, smul(dt, n))), smul(sqrtf(discriminant), n));
        return 1;
    }
    return 0;
}
static vec3 reflect(vec3 v, vec3 n) { return vsub(v, smul(2 * dot(v, n), n)); }

static vec3 material_emitted(const rtnw_scene_desc* d, int mat, float u, float v, vec3 p) {
    const rtnw_material* m = &d->materials[mat];
    if (m->kind == RTNW_MAT_DIFFUSE_LIGHT) return texture_value(d, m->tex, u, v, p);
    return V(0, 0, 0);
}
static int material_scatter(const ctx* c, int mat, const ray* r_in, const hit_record* rec, vec3* attenuation, ray* scattered) {
    const rtnw_scene_desc* d = c->d;
    const rtnw_material* m = &d->materials[mat];
    switch (m->kind) {
        case RTNW_MAT_LAMBERTIAN: {
            vec3 target = vadd(vadd(rec->p, rec->normal), random_in_unit_sphere(c->g));
            scattered->A = rec->p; scattered->B = vsub(target, rec->p); scattered->time = r_in->time;
            *attenuation = texture_value(d, m->tex, rec->u, rec->v, rec->p);
            return 1;
        }
        case RTNW_MAT_METAL: {
            vec3 reflected = reflect(unit_vector(r_in->B), rec->normal);
            scattered->A = rec->p; scattered->B = vadd(reflected, smul(m->f, random_in_unit_sphere(c->g))); scattered->time = 0;
            *attenuation = V(m->albedo[0], m->albedo[1], m->albedo[2]);
            return dot(scattered->B, rec->normal) > 0;
        }
        case RTNW_MAT_DIELECTRIC: {
            const float ref_idx = m->f;
            vec3 outward_normal;
            vec3 reflected = reflect(r_in->B, rec->normal);
            float ni_over_nt;
            *attenuation = V(1.0, 1.0, 1.0);
            vec3 refracted = V(0, 0, 0);
            float reflect_prob;
            float cosine;
            if (dot(r_in->B, rec->normal) > 0) {
                outward_normal = vneg(rec->normal);
                ni_over_nt = ref_idx;
                cosine = dot(r_in->B, rec->normal) / length(r_in->B);
                cosine = sqrtf(1 - ref_idx * ref_idx * (1 - cosine * cosine));
            } else {
                outward_normal = rec->normal;
                ni_over_nt = 1.0 / ref_idx;
                cosine = -dot(r_in->B, rec->normal) / length(r_in->B);
            }
            if (refract(r_in->B, outward_normal, ni_over_nt, &refracted)) reflect_prob = schlick(cosine, ref_idx);
            else reflect_prob = 1.0;
            scattered->A = rec->p; scattered->time = 0;
            if (draw(c->g) < reflect_prob) scattered->B = reflected;
            else scattered->B = refracted;
            return 1;
        }
        case RTNW_MAT_ISOTROPIC: {
            scattered->A = rec->p; scattered->B = random_in_unit_sphere(c->g); scattered->time = 0;
            *attenuation = texture_value(d, m->tex, rec->u, rec->v, rec->p);
            return 1;
        }
        default: return 0;
    }
}

/* ---- camera::get_ray + the jitter of the sample loop: PSC/camera.h:41-56, PSC/main.cpp:305-306 ----------- */
static ray camera_ray(const rtnw_camera* cam, int nx, int ny, int i, int j, stream* g) {
    float u = (float)(i + draw(g)) / (float)(nx);
    float v = (float)(j + draw(g)) / (float)(ny);
    vec3 p;
    do { /* vec3(drand48(), drand48(), 0): right to left, first draw -> y */
        const double dy = draw(g), dx = draw(g);
        p = vsub(smul(2.0, V(dx, dy, 0)), V(1, 1, 0));
    } while (dot(p, p) >= 1.0);
    vec3 rd = smul(cam->lens_radius, p);
    vec3 cu = V(cam->u[0], cam->u[1], cam->u[2]), cv = V(cam->v[0], cam->v[1], cam->v[2]);
    vec3 offset = vadd(smul(rd.e[0], cu), smul(rd.e[1], cv));
    float time = cam->time0 + draw(g) * (cam->time1 - cam->time0);
    vec3 origin = V(cam->origin[0], cam->origin[1], cam->origin[2]);
    vec3 llc = V(cam->lower_left_corner[0], cam->lower_left_corner[1], cam->lower_left_corner[2]);
    vec3 hor = V(cam->horizontal[0], cam->horizontal[1], cam->horizontal[2]);
    vec3 ver = V(cam->vertical[0], cam->vertical[1], cam->vertical[2]);
    ray r;
    r.A = vadd(origin, offset);
    r.B = vsub(vsub(vadd(vadd(llc, smul(u, hor)), smul(v, ver)), origin), offset);
    r.time = time;
    return r;
}

/* ---- color(), PSC/main.cpp:25-46, with the chapter snapshots' variations ---------------- */
static vec3 color(const ctx* c, const ray* r, int depth, const rtnw_render_params* P) {
    hit_record rec;
    if (world_hit(c, r, P->t_min, P->t_max, &rec)) {
        ray scattered;
        vec3 attenuation;
        const int emit = (P->flags & RTNW_F_EMIT) != 0;
        vec3 emitted = emit ? material_emitted(c->d, rec.mat, rec.u, rec.v, rec.p) : V(0, 0, 0);
        if (depth < P->max_depth && material_scatter(c, rec.mat, r, &rec, &attenuation, &scattered)) {
            vec3 rest = color(c, &scattered, depth + 1, P);
            if (emit) return vadd(emitted, vmul(attenuation, rest));
            return vmul(attenuation, rest);
        }
        return emitted;
    }
    if (P->background == RTNW_BG_SKY) { /* TNW/Chapter01_Motion Blur.cpp:29-31 */
        vec3 unit_direction = unit_vector(r->B);
        float t = 0.5 * (unit_direction.e[1] + 1.0);
        return vadd(smul(1.0 - t, V(1.0, 1.0, 1.0)), smul(t, V(0.5, 0.7, 1.0)));
    }
    return V(0, 0, 0);
}

/* ================================================================================================ C ABI */
static void stream_init(stream* g, uint64_t seed) {
    memset(g, 0, sizeof *g);
    g->key[0] = (uint32_t)seed;
    g->key[1] = (uint32_t)(seed >> 32);
}
static ray to_ray(const rtnw_ray* q) {
    ray r;
    r.A = V(q->origin[0], q->origin[1], q->origin[2]);
    r.B = V(q->direction[0], q->direction[1], q->direction[2]);
    r.time = q->time;
    return r;
}

int rtnw_oracle_abi_version(void) { return RTNW_ABI_VERSION; }

/* one world->hit() per ray; media draw keyed numbers with pixel slot = ray.key, sample 0, depth 0 */
int rtnw_oracle_trace(const rtnw_scene_desc* d, const rtnw_ray* rays, size_t n, float t_min, float t_max, uint64_t seed, rtnw_hit* out) {
    stream g;
    stream_init(&g, seed);
    ctx c = {d, &g};
    for (size_t i = 0; i < n; ++i) {
        begin_path(&g, rays[i].key, 0);
        const ray r = to_ray(&rays[i]);
        hit_record rec;
        rtnw_hit* h = &out[i];
        memset(h, 0, sizeof *h);
        h->prim_id = -1;
        h->mat_id = -1;
        if (world_hit(&c, &r, t_min, t_max, &rec)) {
            h->prim_id = d->prim_ids ? d->prim_ids[rec.leaf] : rec.leaf;
            h->sub_id = rec.face;
            h->t = rec.t;
            for (int k = 0; k < 3; ++k) { h->p[k] = rec.p.e[k]; h->normal[k] = rec.normal.e[k]; }
            h->u = rec.u; h->v = rec.v;
            h->mat_id = rec.mat;
        }
    }
    return RTNW_OK;
}

/* per-ray work counters of the reference traversal (design aid): counts[2*i] = aabb tests, counts[2*i+1] = primitive tests */
int rtnw_oracle_trace_counts(const rtnw_scene_desc* d, const rtnw_ray* rays, size_t n, float t_min, float t_max, uint64_t seed,
                             uint32_t* counts) {
    stream g;
    stream_init(&g, seed);
    ctx c = {d, &g};
    for (size_t i = 0; i < n; ++i) {
        begin_path(&g, rays[i].key, 0);
        const ray r = to_ray(&rays[i]);
        hit_record rec;
        const uint64_t b0 = g.box_tests, p0 = g.prim_tests;
        world_hit(&c, &r, t_min, t_max, &rec);
        counts[2 * i] = (uint32_t)(g.box_tests - b0);
        counts[2 * i + 1] = (uint32_t)(g.prim_tests - p0);
    }
    return RTNW_OK;
}

/* the sample loop, PSC/main.cpp:299-313; accum = per-pixel sums; stats = {paths, rays, box tests, prim tests, seconds} */
int rtnw_oracle_render(const rtnw_scene_desc* d, const rtnw_camera* cam, const rtnw_render_params* P, float* accum, double* stats) {
    stream g;
    stream_init(&g, P->seed);
    ctx c = {d, &g};
    struct timespec t0, t1;
    uint64_t paths = 0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int j = P->ny - 1; j >= 0; j--) {
        for (int i = 0; i < P->nx; i++) {
            if (P->pixel_count > 0) { /* pixel subset begin + k*stride, k < count */
                const int rel = j * P->nx + i - P->pixel_begin;
                if (rel < 0 || rel % P->pixel_stride != 0 || rel / P->pixel_stride >= P->pixel_count) continue;
            }
            vec3 col = V(0, 0, 0);
            int s_begin = P->sample_begin, s_count = P->sample_count;
            if (P->flags & RTNW_F_ROTATE_SAMPLES) { /* samples (begin - p) mod G + k*G below ns = sample_count */
                const int g_ = P->sample_stride, pix = j * P->nx + i;
                s_begin = ((P->sample_begin - pix) % g_ + g_) % g_;
                s_count = s_begin < P->sample_count ? (P->sample_count - s_begin + g_ - 1) / g_ : 0;
            }
            paths += (uint64_t)s_count;
            for (int k = 0; k < s_count; k++) {
                const int s = s_begin + k * P->sample_stride;
                begin_path(&g, (uint32_t)(j * P->nx + i), (uint32_t)s);
                ray r = camera_ray(cam, P->nx, P->ny, i, j, &g);
                vec3 temp = color(&c, &r, 0, P);
                if (P->flags & RTNW_F_DE_NAN) /* PSC/main.cpp:232-242 */
                    for (int q = 0; q < 3; ++q)
                        if (!(temp.e[q] == temp.e[q])) temp.e[q] = 0;
                col = vadd(col, temp);
            }
            for (int q = 0; q < 3; ++q) {
                float* dst = &accum[3 * ((size_t)j * P->nx + i) + q];
                *dst = (P->flags & RTNW_F_ACCUMULATE) ? *dst + col.e[q] : col.e[q];
            }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (stats) {
        stats[0] = (double)paths;
        stats[1] = (double)g.rays;
        stats[2] = (double)g.box_tests;
        stats[3] = (double)g.prim_tests;
        stats[4] = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    }
    return RTNW_OK;
}

int rtnw_oracle_camera_rays(const rtnw_camera* cam, int32_t nx, int32_t ny, const int32_t* ij, const int32_t* sample, size_t n,
                            uint64_t seed, rtnw_ray* out) {
    stream g;
    stream_init(&g, seed);
    for (size_t q = 0; q < n; ++q) {
        const int i = ij[2 * q], j = ij[2 * q + 1];
        begin_path(&g, (uint32_t)(j * nx + i), (uint32_t)sample[q]);
        const ray r = camera_ray(cam, nx, ny, i, j, &g);
        for (int k = 0; k < 3; ++k) { out[q].origin[k] = r.A.e[k]; out[q].direction[k] = r.B.e[k]; }
        out[q].time = r.time;
        out[q].key = (uint32_t)(j * nx + i);
    }
    return RTNW_OK;
}

int rtnw_oracle_eval_texture(const rtnw_scene_desc* d, int32_t tex, const float* uvp, size_t n, float* rgb) {
    if (tex < 0 || tex >= d->n_textures) return RTNW_ERR_INVALID;
    for (size_t i = 0; i < n; ++i) {
        const float* q = uvp + 5 * i;
        const vec3 v = texture_value(d, tex, q[0], q[1], V(q[2], q[3], q[4]));
        rgb[3 * i] = v.e[0]; rgb[3 * i + 1] = v.e[1]; rgb[3 * i + 2] = v.e[2];
    }
    return RTNW_OK;
}

int rtnw_oracle_eval_perlin(const rtnw_scene_desc* d, int32_t which, const float* xyz, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i) {
        const vec3 p = V(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        out[i] = which == 0 ? perlin_noise(d, p) : perlin_turb(d, p);
    }
    return RTNW_OK;
}

/* material::emitted + material::scatter; draws from the sequential stream (seed, pixel = i, sample 0) */
int rtnw_oracle_scatter(const rtnw_scene_desc* d, const rtnw_ray* rays_in, const rtnw_hit* hits, size_t n, uint64_t seed,
                        rtnw_ray* out_sc, float* out_att, float* out_em, int32_t* out_flag) {
    stream g;
    stream_init(&g, seed);
    ctx c = {d, &g};
    for (size_t i = 0; i < n; ++i) {
        if (hits[i].mat_id < 0 || hits[i].mat_id >= d->n_materials) return RTNW_ERR_INVALID;
        begin_path(&g, (uint32_t)i, 0);
        const ray r = to_ray(&rays_in[i]);
        hit_record rec;
        memset(&rec, 0, sizeof rec);
        rec.t = hits[i].t; rec.u = hits[i].u; rec.v = hits[i].v;
        rec.p = V(hits[i].p[0], hits[i].p[1], hits[i].p[2]);
        rec.normal = V(hits[i].normal[0], hits[i].normal[1], hits[i].normal[2]);
        rec.mat = hits[i].mat_id;
        const vec3 em = material_emitted(d, rec.mat, rec.u, rec.v, rec.p);
        vec3 att = V(0, 0, 0);
        ray sc;
        memset(&sc, 0, sizeof sc);
        const int ok = material_scatter(&c, rec.mat, &r, &rec, &att, &sc);
        memset(&out_sc[i], 0, sizeof out_sc[i]);
        for (int k = 0; k < 3; ++k) {
            out_em[3 * i + k] = em.e[k];
            out_att[3 * i + k] = ok ? att.e[k] : 0.0f;
            out_sc[i].origin[k] = ok ? sc.A.e[k] : 0.0f;
            out_sc[i].direction[k] = ok ? sc.B.e[k] : 0.0f;
        }
        out_sc[i].time = ok ? sc.time : 0.0f;
        out_sc[i].key = (uint32_t)i;
        out_flag[i] = ok;
    }
    return RTNW_OK;
}
