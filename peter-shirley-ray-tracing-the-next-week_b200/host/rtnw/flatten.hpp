// Scene flattener: host object graph (scene.hpp) -> SoA tables of include/rtnw.h.
//
// Canonical form: the top-level hitable_list becomes an ordered sequence of ITEMS; nested lists are
// inlined (list-in-list has the same narrowing semantics as one list, PSC/hitable_list.h:20-32); translate /
// rotate_y wrappers become per-primitive or per-item transform chains; flip_normals becomes a parity bit on the
// primitives beneath it; every maximal tree of bvh_node objects becomes one BVH item whose leaves are primitive
// ranges; constant_medium becomes a primitive that points at a range of boundary primitives.
#ifndef RTNW_FLATTEN_HPP_
#define RTNW_FLATTEN_HPP_

#include <cstdint>
#include <string>
#include <vector>

#include "rtnw.h"
#include "rtnw/scene.hpp"

namespace rtnw {

struct flat_scene {
    std::vector<rtnw_item> items;
    std::vector<rtnw_bvh_node> nodes;
    std::vector<rtnw_prim> prims;
    std::vector<int32_t> prim_ids;
    std::vector<rtnw_xform_op> xforms;
    std::vector<rtnw_material> materials;
    std::vector<rtnw_texture> textures;
    std::vector<uint8_t> images;
    float ranvec[256 * 3];
    int32_t perm_x[256], perm_y[256], perm_z[256];
    // index -> host object, for the device bridge and for diagnostics
    std::vector<const material*> material_objects;
    std::vector<const texture*> texture_objects;
    int32_t n_leaves = 0;
    std::string error;

    rtnw_scene_desc desc() const;
};

// Returns RTNW_OK, or RTNW_ERR_UNSUPPORTED / RTNW_ERR_INVALID with out.error set.
int flatten(const hitable* world, flat_scene& out);

void to_c_camera(const camera& cam, rtnw_camera& out);

}  // namespace rtnw

#endif
