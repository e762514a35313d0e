// drop-in for the reference header of the same name (PSC/vec3.h): every class lives in rtnw/scene.hpp
#pragma once
#include "rtnw/scene.hpp"
