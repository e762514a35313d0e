// Chapter scene builders — the host-side users of the scene API that define BASELINE.json's five configs
// (SURVEY.md §8d).  They are written against rtnw/scene.hpp exactly as the reference's builders are written
// against its headers (raw `new`, a hitable** array, `return new hitable_list(list, n)`), and they consume
// drand48 in the order the reference binary (g++, x86-64) does, so the scenes are object-for-object identical
// to the reference's.
#ifndef RTNW_CHAPTER_SCENES_HPP_
#define RTNW_CHAPTER_SCENES_HPP_

#include "rtnw/scene.hpp"

namespace rtnw_scenes {

struct view {  // the camera + integrator settings each chapter snapshot hard-codes in its main()
    vec3 lookfrom, lookat;
    float vfov, aperture, focus_dist, time0, time1;
    int nx, ny, ns;          // the config's image size and samples per pixel
    float t_min;             // `color`'s t_min (PSC/main.cpp:27; 0.0 in Ch01/Ch03, 0.01 in Ch07/Ch08)
    bool sky;                // Ch01-Ch04 sky gradient vs black
    bool emit;               // emitted term present
    bool de_nan;             // PSC/main.cpp:311
};

hitable* random_scene_ch01();      // config 1: TNW/Chapter01_Motion Blur.cpp:36-67 with constant_texture albedos
hitable* two_perlin_spheres();     // config 2: checker ground + noise_texture(4) sphere (PSC/main.cpp:112-145 classes)
hitable* cornell_box();            // config 3: PSC/main.cpp:148-166
hitable* cornell_smoke();          // config 4: PSC/main.cpp:169-188
hitable* final_scene();            // config 5R: PSC/main.cpp:190-230 (flat list, nb = 10)
hitable* final_northstar();        // config 5N: nb = 32 floor boxes in a bvh_node, 1000-sphere cluster in
                                   //            translate(rotate_y(bvh_node)), synthetic earth image texture
hitable* random_scene();           // PSC/main.cpp:48-85 (live version: checker ground, metal + glass spheres)
hitable* test_scene();             // PSC/main.cpp:135-145
hitable* simple_light();           // PSC/main.cpp:122-133
hitable* two_spheres();            // PSC/main.cpp:99-110
hitable* earth();                  // PSC/main.cpp:87-97 with the synthetic RGB8 image
hitable* earth(unsigned char* rgb, int nx, int ny);  // ... with a caller-supplied RGB8 image (decoded PNG)
hitable* stress_shells();          // test fixture (not in the reference): 600 nested shells, every ray passes every box
hitable* twin_bvh();               // test fixture: bvh_nodes that are neighbours in the top-level list
hitable* wrap_in_bvh(hitable* flat_list, float t0, float t1);  // `new bvh_node(list->list, list->list_size, t0, t1)`

// deterministic 1024x512 RGB8 stand-in for picture.png (the shipped PNG is RGBA and mis-strided)
unsigned char* synthetic_earth(int& nx, int& ny);

view view_ch01();        // (13,2,3)->(0,0,0), vfov 20, aperture 0.1, 200x100x100, sky
view view_two_perlin();  // same camera, aperture 0, 400x200x256, sky
view view_cornell();     // (278,278,-800)->(278,278,0), vfov 40, 500x500x1000, black
view view_final();       // (228,278,-800)->(278,278,0), vfov 40, 1000x1000x100, black (PSC/main.cpp:248-259)

camera make_camera(const view& v, int nx, int ny);

}  // namespace rtnw_scenes

#endif
