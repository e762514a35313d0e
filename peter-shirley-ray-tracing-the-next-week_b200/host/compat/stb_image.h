// drop-in for the one stb_image call the reference makes (stbi_load("picture.png", &nx, &ny, &nn, 0), PSC/main.cpp:93): decodes
// the PNG with the host library's own decoder into tightly packed RGB8 (*comp = 3), the layout image_texture indexes.
#pragma once
#include "rtnw_host.h"
static inline unsigned char* stbi_load(const char* path, int* x, int* y, int* comp, int /*req_comp*/) {
    unsigned char* px = nullptr;
    int32_t nx = 0, ny = 0;
    if (rtnw_host_load_png(path, &px, &nx, &ny) != RTNW_OK) return nullptr;
    *x = nx; *y = ny; if (comp) *comp = 3;
    return px;
}
