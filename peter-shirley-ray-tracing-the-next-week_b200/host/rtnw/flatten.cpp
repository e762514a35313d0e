#include "rtnw/flatten.hpp"

#include <cstring>
#include <map>
#include <stdexcept>

namespace rtnw {
namespace {

struct unsupported : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline float int_bits(int32_t v) {
    float f;
    std::memcpy(&f, &v, 4);
    return f;
}

typedef std::vector<rtnw_xform_op> chain_t;

struct op_less {
    bool operator()(const chain_t& a, const chain_t& b) const {
        if (a.size() != b.size()) return a.size() < b.size();
        return std::memcmp(a.data(), b.data(), a.size() * sizeof(rtnw_xform_op)) < 0;
    }
};

class flattener {
public:
    explicit flattener(flat_scene& s) : S(s) {}

    void run(const hitable* world) {
        S.xforms.push_back(rtnw_xform_op{0, 0, 0, RTNW_XF_END});  // op 0 = identity chain
        number_leaves(world);
        S.n_leaves = next_leaf_;
        emit_world(world, chain_t(), false);
        // boundary primitives live after every item range; patch the media that point at them
        const int32_t base = (int32_t)S.prims.size();
        S.prims.insert(S.prims.end(), boundary_.begin(), boundary_.end());
        S.prim_ids.insert(S.prim_ids.end(), boundary_.size(), -1);
        for (const fixup& f : fixups_) {
            S.prims[f.medium_slot].f[1] = int_bits(base + f.first);
            S.prims[f.medium_slot].f[2] = int_bits(f.count);
        }
        copy_perlin();
    }

private:
    flat_scene& S;
    std::vector<rtnw_prim> boundary_;
    struct fixup { int32_t medium_slot, first, count; };
    std::vector<fixup> fixups_;
    std::map<const hitable*, int32_t> leaf_id_;
    int32_t next_leaf_ = 0;
    std::map<const material*, int32_t> mat_index_;
    std::map<const texture*, int32_t> tex_index_;
    std::map<const unsigned char*, int32_t> image_offset_;
    std::map<chain_t, uint32_t, op_less> chain_index_;
    bool uses_noise_ = false;

    // ---- leaf ids: creation-order index of each leaf handed to the list / BVH ----
    void number_leaves(const hitable* h) {
        switch (h->rtnw_kind()) {
            case geo_kind::list: {
                const hitable_list* l = static_cast<const hitable_list*>(h);
                for (int i = 0; i < l->list_size; ++i) number_leaves(l->list[i]);
                break;
            }
            case geo_kind::bvh: {
                const bvh_node* n = static_cast<const bvh_node*>(h);
                if (!n->creation_order.empty()) {
                    for (const hitable* c : n->creation_order) number_leaves(c);
                } else {  // inner node reached directly (user kept a pointer to a subtree)
                    number_leaves(n->left);
                    if (n->right != n->left) number_leaves(n->right);
                }
                break;
            }
            case geo_kind::flip: number_leaves(static_cast<const flip_normals*>(h)->ptr); break;
            case geo_kind::translate: number_leaves(static_cast<const translate*>(h)->ptr); break;
            case geo_kind::rotate_y: number_leaves(static_cast<const rotate_y*>(h)->ptr); break;
            case geo_kind::user: throw unsupported("user-defined hitable subclass cannot be flattened");
            default:
                if (!leaf_id_.count(h)) leaf_id_[h] = next_leaf_++;
        }
    }

    // ---- tables ----
    uint32_t intern_chain(const chain_t& ops) {
        if (ops.empty()) return 0;
        auto it = chain_index_.find(ops);
        if (it != chain_index_.end()) return it->second;
        const uint32_t first = (uint32_t)S.xforms.size();
        if (first + ops.size() >= (1u << 28)) throw unsupported("too many transform ops");
        for (size_t i = 0; i < ops.size(); ++i) S.xforms.push_back(ops[i]);
        S.xforms[first].kind |= (uint32_t)ops.size() << 8;
        chain_index_[ops] = first;
        return first;
    }

    int32_t intern_texture(const texture* t) {
        if (!t) throw unsupported("null texture");
        auto it = tex_index_.find(t);
        if (it != tex_index_.end()) return it->second;
        rtnw_texture rec;
        std::memset(&rec, 0, sizeof rec);
        // reserve the slot first so recursive checkers keep parent-before-child order
        const int32_t idx = (int32_t)S.textures.size();
        S.textures.push_back(rec);
        S.texture_objects.push_back(t);
        tex_index_[t] = idx;
        switch (t->rtnw_kind()) {
            case tex_kind::constant: {
                const vec3& c = static_cast<const constant_texture*>(t)->color;
                rec.kind = RTNW_TEX_CONSTANT;
                rec.c[0] = c.e[0]; rec.c[1] = c.e[1]; rec.c[2] = c.e[2];
                break;
            }
            case tex_kind::checker: {
                const checker_texture* c = static_cast<const checker_texture*>(t);
                rec.kind = RTNW_TEX_CHECKER;
                rec.i0 = intern_texture(c->even);
                rec.i1 = intern_texture(c->odd);
                break;
            }
            case tex_kind::noise:
                rec.kind = RTNW_TEX_NOISE;
                rec.c[0] = static_cast<const noise_texture*>(t)->scale;
                uses_noise_ = true;
                break;
            case tex_kind::readme_noise: {
                const int variant = static_cast<const readme_noise_texture*>(t)->variant;
                if (variant < 1 || variant > 3) throw unsupported("readme_noise_texture variant must be 1, 2 or 3");
                rec.kind = RTNW_TEX_NOISE_HASH + (variant - 1);
                uses_noise_ = true;
                break;
            }
            case tex_kind::image: {
                const image_texture* im = static_cast<const image_texture*>(t);
                if (!im->data || im->nx <= 0 || im->ny <= 0) throw unsupported("image_texture without pixels");
                rec.kind = RTNW_TEX_IMAGE;
                auto found = image_offset_.find(im->data);
                if (found == image_offset_.end()) {
                    while (S.images.size() % 16) S.images.push_back(0);
                    const size_t off = S.images.size();
                    if (off + (size_t)im->nx * im->ny * 3 > 0x7fffffffu) throw unsupported("image pool exceeds 2 GiB");
                    S.images.insert(S.images.end(), im->data, im->data + (size_t)im->nx * im->ny * 3);
                    found = image_offset_.insert(std::make_pair(im->data, (int32_t)off)).first;
                }
                rec.flags = im->bilinear ? RTNW_TEXF_BILINEAR : 0u;
                rec.i0 = found->second;
                rec.i1 = im->nx;
                rec.i2 = im->ny;
                break;
            }
            default: throw unsupported("user-defined texture subclass cannot be flattened");
        }
        S.textures[idx] = rec;
        return idx;
    }

    int32_t intern_material(const material* m) {
        if (!m) throw unsupported("null material");
        auto it = mat_index_.find(m);
        if (it != mat_index_.end()) return it->second;
        rtnw_material rec;
        std::memset(&rec, 0, sizeof rec);
        rec.tex = -1;
        switch (m->rtnw_kind()) {
            case mat_kind::lambertian:
                rec.kind = RTNW_MAT_LAMBERTIAN;
                rec.tex = intern_texture(static_cast<const lambertian*>(m)->albedo);
                break;
            case mat_kind::metal: {
                const metal* mm = static_cast<const metal*>(m);
                rec.kind = RTNW_MAT_METAL;
                rec.f = mm->fuzz;
                rec.albedo[0] = mm->albedo.e[0]; rec.albedo[1] = mm->albedo.e[1]; rec.albedo[2] = mm->albedo.e[2];
                break;
            }
            case mat_kind::dielectric:
                rec.kind = RTNW_MAT_DIELECTRIC;
                rec.f = static_cast<const dielectric*>(m)->ref_idx;
                break;
            case mat_kind::diffuse_light:
                rec.kind = RTNW_MAT_DIFFUSE_LIGHT;
                rec.tex = intern_texture(static_cast<const diffuse_light*>(m)->emit);
                break;
            case mat_kind::isotropic:
                rec.kind = RTNW_MAT_ISOTROPIC;
                rec.tex = intern_texture(static_cast<const isotropic*>(m)->albedo);
                break;
            default: throw unsupported("user-defined material subclass cannot be flattened");
        }
        const int32_t idx = (int32_t)S.materials.size();
        S.materials.push_back(rec);
        S.material_objects.push_back(m);
        mat_index_[m] = idx;
        return idx;
    }

    static rtnw_prim blank(uint32_t kind, bool flip, uint32_t xform, int32_t mat) {
        rtnw_prim p;
        std::memset(&p, 0, sizeof p);
        p.kx = RTNW_KX(kind, flip ? 1 : 0, xform);
        p.mat = mat;
        return p;
    }

    // Append the primitives of `h` (anything that is not / does not contain a BVH) to dst with list semantics.
    // `ids` is null when dst is the boundary table.
    void emit_leaves(const hitable* h, const chain_t& chain, bool flip, std::vector<rtnw_prim>& dst,
                     std::vector<int32_t>* ids) {
        const int32_t id = (ids && leaf_id_.count(h)) ? leaf_id_[h] : -1;
        auto push = [&](const rtnw_prim& p) {
            dst.push_back(p);
            if (ids) ids->push_back(id);
        };
        switch (h->rtnw_kind()) {
            case geo_kind::sphere: {
                const sphere* s = static_cast<const sphere*>(h);
                rtnw_prim p = blank(RTNW_PRIM_SPHERE, flip, intern_chain(chain), ids ? intern_material(s->mat_ptr) : -1);
                p.f[0] = s->center.e[0]; p.f[1] = s->center.e[1]; p.f[2] = s->center.e[2]; p.f[3] = s->radius;
                push(p);
                break;
            }
            case geo_kind::moving_sphere: {
                const moving_sphere* s = static_cast<const moving_sphere*>(h);
                rtnw_prim p = blank(RTNW_PRIM_MOVING_SPHERE, flip, intern_chain(chain), ids ? intern_material(s->mat_ptr) : -1);
                p.f[0] = s->center0.e[0]; p.f[1] = s->center0.e[1]; p.f[2] = s->center0.e[2]; p.f[3] = s->radius;
                p.f[4] = s->time0; p.f[5] = s->time1;
                push(p);
                rtnw_prim e = blank(RTNW_PRIM_EXT, false, 0, -1);
                e.f[0] = s->center1.e[0]; e.f[1] = s->center1.e[1]; e.f[2] = s->center1.e[2];
                push(e);
                break;
            }
            case geo_kind::rect_xy: emit_rect(static_cast<const xy_rect*>(h), RTNW_PRIM_RECT_XY, chain, flip, push, ids != nullptr); break;
            case geo_kind::rect_xz: emit_rect(static_cast<const xz_rect*>(h), RTNW_PRIM_RECT_XZ, chain, flip, push, ids != nullptr); break;
            case geo_kind::rect_yz: emit_rect(static_cast<const yz_rect*>(h), RTNW_PRIM_RECT_YZ, chain, flip, push, ids != nullptr); break;
            case geo_kind::box: {
                const box* b = static_cast<const box*>(h);
                rtnw_prim p = blank(RTNW_PRIM_BOX, flip, intern_chain(chain), ids ? intern_material(b->mat_ptr) : -1);
                for (int c = 0; c < 3; ++c) { p.f[c] = b->pmin.e[c]; p.f[3 + c] = b->pmax.e[c]; }
                push(p);
                break;
            }
            case geo_kind::medium: {
                if (!ids) throw unsupported("constant_medium used as the boundary of another constant_medium");
                const constant_medium* m = static_cast<const constant_medium*>(h);
                rtnw_prim p = blank(RTNW_PRIM_MEDIUM, flip, intern_chain(chain), intern_material(m->phase_function));
                p.f[0] = m->density;
                p.f[3] = int_bits(id);  // keys the free-flight draw
                const int32_t first = (int32_t)boundary_.size();
                emit_leaves(m->boundary, chain_t(), false, boundary_, nullptr);
                fixups_.push_back(fixup{(int32_t)dst.size(), first, (int32_t)boundary_.size() - first});
                push(p);
                break;
            }
            case geo_kind::list: {
                const hitable_list* l = static_cast<const hitable_list*>(h);
                for (int i = 0; i < l->list_size; ++i) emit_leaves(l->list[i], chain, flip, dst, ids);
                break;
            }
            case geo_kind::flip: emit_leaves(static_cast<const flip_normals*>(h)->ptr, chain, !flip, dst, ids); break;
            case geo_kind::translate: {
                const translate* t = static_cast<const translate*>(h);
                chain_t c2(chain);
                c2.push_back(rtnw_xform_op{t->offset.e[0], t->offset.e[1], t->offset.e[2], RTNW_XF_TRANSLATE});
                emit_leaves(t->ptr, c2, flip, dst, ids);
                break;
            }
            case geo_kind::rotate_y: {
                const rotate_y* r = static_cast<const rotate_y*>(h);
                chain_t c2(chain);
                c2.push_back(rtnw_xform_op{r->sin_theta, r->cos_theta, 0, RTNW_XF_ROTATE_Y});
                emit_leaves(r->ptr, c2, flip, dst, ids);
                break;
            }
            case geo_kind::bvh:
                throw unsupported("bvh_node nested below a BVH leaf, a medium boundary, or a list inside a BVH");
            default: throw unsupported("user-defined hitable subclass cannot be flattened");
        }
    }

    template <class R, class Push>
    void emit_rect(const R* r, uint32_t kind, const chain_t& chain, bool flip, Push& push, bool with_mat) {
        rtnw_prim p = blank(kind, flip, intern_chain(chain), with_mat ? intern_material(r->mp) : -1);
        p.f[0] = r->a0; p.f[1] = r->a1; p.f[2] = r->b0; p.f[3] = r->b1; p.f[4] = r->k;
        push(p);
    }

    // child of a bvh_node -> (ref, count, box)
    void emit_child(const hitable* c, const bvh_node* parent, bool flip, int32_t& ref, int32_t& count, float* bmin, float* bmax) {
        aabb b;
        if (c->rtnw_kind() == geo_kind::bvh) {
            const bvh_node* n = static_cast<const bvh_node*>(c);
            b = n->box;
            ref = emit_tree(n, flip);
            count = 0;
        } else {
            if (!c->bounding_box(parent->time0, parent->time1, b)) throw unsupported("BVH leaf without a bounding box");
            const int32_t first = (int32_t)S.prims.size();
            emit_leaves(c, chain_t(), flip, S.prims, &S.prim_ids);
            ref = ~first;
            count = (int32_t)S.prims.size() - first;
        }
        for (int a = 0; a < 3; ++a) { bmin[a] = b.min().e[a]; bmax[a] = b.max().e[a]; }
    }

    int32_t emit_tree(const bvh_node* n, bool flip) {
        const int32_t idx = (int32_t)S.nodes.size();
        S.nodes.push_back(rtnw_bvh_node());
        rtnw_bvh_node rec;
        std::memset(&rec, 0, sizeof rec);
        emit_child(n->left, n, flip, rec.left, rec.lcount, rec.lmin, rec.lmax);
        if (n->right == n->left) {
            rec.right = RTNW_REF_NONE;  // the reference tests the single leaf twice; the result is the same leaf
            rec.rcount = 0;
            std::memcpy(rec.rmin, rec.lmin, sizeof rec.rmin);
            std::memcpy(rec.rmax, rec.lmax, sizeof rec.rmax);
        } else {
            emit_child(n->right, n, flip, rec.right, rec.rcount, rec.rmin, rec.rmax);
        }
        S.nodes[idx] = rec;
        return idx;
    }

    void emit_world(const hitable* h, const chain_t& chain, bool flip) {
        switch (h->rtnw_kind()) {
            case geo_kind::list: {
                const hitable_list* l = static_cast<const hitable_list*>(h);
                for (int i = 0; i < l->list_size; ++i) emit_world(l->list[i], chain, flip);
                break;
            }
            case geo_kind::flip: emit_world(static_cast<const flip_normals*>(h)->ptr, chain, !flip); break;
            case geo_kind::translate: {
                const translate* t = static_cast<const translate*>(h);
                chain_t c2(chain);
                c2.push_back(rtnw_xform_op{t->offset.e[0], t->offset.e[1], t->offset.e[2], RTNW_XF_TRANSLATE});
                emit_world(t->ptr, c2, flip);
                break;
            }
            case geo_kind::rotate_y: {
                const rotate_y* r = static_cast<const rotate_y*>(h);
                chain_t c2(chain);
                c2.push_back(rtnw_xform_op{r->sin_theta, r->cos_theta, 0, RTNW_XF_ROTATE_Y});
                emit_world(r->ptr, c2, flip);
                break;
            }
            case geo_kind::bvh: {
                const bvh_node* n = static_cast<const bvh_node*>(h);
                rtnw_item it;
                std::memset(&it, 0, sizeof it);
                it.kind = RTNW_ITEM_BVH;
                it.xform = intern_chain(chain);
                it.flip = flip ? 1 : 0;
                for (int a = 0; a < 3; ++a) { it.bmin[a] = n->box.min().e[a]; it.bmax[a] = n->box.max().e[a]; }
                const int32_t before = (int32_t)S.nodes.size();
                it.first = emit_tree(n, flip);
                it.count = (int32_t)S.nodes.size() - before;
                S.items.push_back(it);
                break;
            }
            case geo_kind::user: throw unsupported("user-defined hitable subclass cannot be flattened");
            default: {  // a leaf directly in the world list: extend the current run of primitives
                const int32_t first = (int32_t)S.prims.size();
                emit_leaves(h, chain, flip, S.prims, &S.prim_ids);
                const int32_t added = (int32_t)S.prims.size() - first;
                if (!S.items.empty() && S.items.back().kind == RTNW_ITEM_PRIMS &&
                    S.items.back().first + S.items.back().count == first) {
                    S.items.back().count += added;
                } else {
                    rtnw_item it;
                    std::memset(&it, 0, sizeof it);
                    it.kind = RTNW_ITEM_PRIMS;
                    it.first = first;
                    it.count = added;
                    S.items.push_back(it);
                }
            }
        }
    }

    void copy_perlin() {
        if (!perlin::ranvec) perlin::regenerate();
        for (int i = 0; i < 256; ++i) {
            for (int c = 0; c < 3; ++c) S.ranvec[3 * i + c] = perlin::ranvec[i].e[c];
            S.perm_x[i] = perlin::perm_x[i];
            S.perm_y[i] = perlin::perm_y[i];
            S.perm_z[i] = perlin::perm_z[i];
        }
    }
};

}  // namespace

rtnw_scene_desc flat_scene::desc() const {
    rtnw_scene_desc d;
    std::memset(&d, 0, sizeof d);
    d.abi_version = RTNW_ABI_VERSION;
    d.n_items = (int32_t)items.size();        d.items = items.data();
    d.n_nodes = (int32_t)nodes.size();        d.nodes = nodes.data();
    d.n_prim_slots = (int32_t)prims.size();   d.prims = prims.data();
    d.prim_ids = prim_ids.data();
    d.n_xform_ops = (int32_t)xforms.size();   d.xforms = xforms.data();
    d.n_materials = (int32_t)materials.size(); d.materials = materials.data();
    d.n_textures = (int32_t)textures.size();  d.textures = textures.data();
    d.image_bytes = images.size();            d.images = images.data();
    d.perlin_ranvec = ranvec;
    d.perlin_perm_x = perm_x;
    d.perlin_perm_y = perm_y;
    d.perlin_perm_z = perm_z;
    return d;
}

int flatten(const hitable* world, flat_scene& out) {
    out = flat_scene();
    if (!world) {
        out.error = "null world";
        return RTNW_ERR_INVALID;
    }
    try {
        flattener f(out);
        f.run(world);
    } catch (const unsupported& e) {
        out.error = e.what();
        return RTNW_ERR_UNSUPPORTED;
    } catch (const std::exception& e) {
        out.error = e.what();
        return RTNW_ERR_INVALID;
    }
    return RTNW_OK;
}

void to_c_camera(const camera& cam, rtnw_camera& out) {
    for (int c = 0; c < 3; ++c) {
        out.origin[c] = cam.origin.e[c];
        out.lower_left_corner[c] = cam.lower_left_corner.e[c];
        out.horizontal[c] = cam.horizontal.e[c];
        out.vertical[c] = cam.vertical.e[c];
        out.u[c] = cam.u.e[c];
        out.v[c] = cam.v.e[c];
        out.w[c] = cam.w.e[c];
    }
    out.lens_radius = cam.len_radius;
    out.time0 = cam.time0;
    out.time1 = cam.time1;
}

}  // namespace rtnw
