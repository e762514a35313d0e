// png.cpp — texture ingest for image_texture: decode a PNG into the tightly packed RGB8 array that
// image_texture::value indexes (PSC/surface_texture.h:19-30: data[3*i + 3*nx*j + c]).
//
// Replaces the reference's `stbi_load("picture.png", &nx, &ny, &nn, 0)` (PSC/main.cpp:93, stb_image v2.06 vendored in
// the reference tree).  The reference passes stb's buffer on with whatever channel count the file has while
// image_texture assumes three (a 4-channel file is indexed with stride 3); here the alpha channel is dropped
// and grey / palette images are expanded, so the texture table always holds RGB8.
// Scope: 8-bit and 16-bit samples (16-bit keeps the high byte), colour types 0/2/3/4/6, non-interlaced; zlib does the
// inflate.  Written from the PNG specification (filter types 0-4, Paeth predictor), not from stb_image.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace rtnw {

namespace {
uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]; }
int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

// Returns a new[]-allocated nx*ny*3 array (the scene API's image_texture keeps the pointer, as the reference does with
// stb's), or nullptr with `err` set.
unsigned char* load_png_rgb8(const char* path, int& nx, int& ny, std::string& err) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return nullptr; }
    std::vector<unsigned char> file;
    unsigned char buf[65536];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + got);
    std::fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) { err = "not a PNG file"; return nullptr; }
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<unsigned char> idat, palette;
    size_t at = 8;
    bool end = false;
    while (!end && at + 12 <= file.size()) {
        const uint32_t len = be32(&file[at]);
        const unsigned char* type = &file[at + 4];
        if ((size_t)len > file.size() - at - 12) { err = "truncated PNG chunk"; return nullptr; }
        const unsigned char* data = &file[at + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad IHDR"; return nullptr; }
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        at += 12 + (size_t)len;
    }
    if (ctype < 0 || w == 0 || h == 0 || w > 32768 || h > 32768) { err = "missing or implausible IHDR"; return nullptr; }
    if (interlace) { err = "interlaced PNG is not supported"; return nullptr; }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "unknown PNG colour type"; return nullptr;
    }
    if (!(depth == 8 || (depth == 16 && ctype != 3))) { err = "only 8-bit (and 16-bit non-palette) PNG samples are supported"; return nullptr; }
    if (ctype == 3 && palette.size() < 3) { err = "palette PNG without PLTE"; return nullptr; }
    const size_t bpp = (size_t)channels * (depth / 8), stride = bpp * w;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf raw_len = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || raw_len != raw.size()) { err = "PNG image data does not inflate to the size IHDR announces"; return nullptr; }
    // undo the per-row filters in place (row r starts at r*(stride+1) with its filter byte)
    std::vector<unsigned char> prev(stride, 0);
    for (uint32_t r = 0; r < h; ++r) {
        unsigned char* row = &raw[(size_t)r * (stride + 1)];
        const int ft = row[0];
        unsigned char* x = row + 1;
        if (ft > 4) { err = "bad PNG filter type"; return nullptr; }
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? x[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int add = 0;
            if (ft == 1) add = a; else if (ft == 2) add = b; else if (ft == 3) add = (a + b) >> 1; else if (ft == 4) add = paeth(a, b, c);
            x[i] = (unsigned char)(x[i] + add);
        }
        std::memcpy(prev.data(), x, stride);
    }
    nx = (int)w;
    ny = (int)h;
    unsigned char* rgb = new unsigned char[(size_t)w * h * 3];
    const size_t sample = depth / 8;  // a 16-bit sample keeps its high (first) byte
    for (uint32_t r = 0; r < h; ++r) {
        const unsigned char* x = &raw[(size_t)r * (stride + 1) + 1];
        unsigned char* o = rgb + (size_t)r * w * 3;
        for (uint32_t i = 0; i < w; ++i) {
            const unsigned char* p = x + (size_t)i * bpp;
            if (ctype == 2 || ctype == 6) { o[3 * i] = p[0]; o[3 * i + 1] = p[sample]; o[3 * i + 2] = p[2 * sample]; }
            else if (ctype == 3) {
                const size_t e = 3 * (size_t)p[0];
                if (e + 3 > palette.size()) { delete[] rgb; err = "palette index out of range"; return nullptr; }
                o[3 * i] = palette[e]; o[3 * i + 1] = palette[e + 1]; o[3 * i + 2] = palette[e + 2];
            } else { o[3 * i] = o[3 * i + 1] = o[3 * i + 2] = p[0]; }
        }
    }
    return rgb;
}

}  // namespace rtnw
