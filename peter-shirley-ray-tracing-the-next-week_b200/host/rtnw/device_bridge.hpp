// The reference API has four virtuals that do arithmetic on the hot path (hitable::hit, material::scatter,
// material::emitted, texture::value; PSC/hitable.h:34, PSC/material.h:54-57, PSC/texture.h:13) and camera::get_ray
// (PSC/camera.h:41-47).  In this
// framework they have no CPU implementation: a bridge forwards single calls to the GPU library
// (rtnw_trace / rtnw_scatter / rtnw_eval_texture).  Without an installed bridge the calls abort loudly.
#ifndef RTNW_DEVICE_BRIDGE_HPP_
#define RTNW_DEVICE_BRIDGE_HPP_

#include "rtnw/scene.hpp"

namespace rtnw {
struct device_bridge {
    bool (*hit)(const hitable*, const ray&, float, float, hit_record&) = nullptr;
    bool (*scatter)(const material*, const ray&, const hit_record&, vec3&, ray&) = nullptr;
    vec3 (*emitted)(const material*, float, float, const vec3&) = nullptr;
    vec3 (*value)(const texture*, float, float, const vec3&) = nullptr;
    ray (*get_ray)(const camera*, float, float) = nullptr;
};
void set_device_bridge(const device_bridge& b);
const device_bridge& get_device_bridge();
// defined in cuda_bridge.cpp (programs that link librtnw.so): serve the four virtuals from GPU `device`
int install_cuda_bridge(int device, unsigned long long seed = 1);
// the bridge caches the device form of every object a call was made on (by address): worlds must not change after their first
// call, or be dropped with bridge_invalidate(object); bridge_release() frees every cached scene and the bridge's context
void bridge_invalidate(const void* object);
void bridge_release();
}  // namespace rtnw

#endif
