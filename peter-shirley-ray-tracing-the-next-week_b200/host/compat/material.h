// drop-in for the reference header of the same name (PSC/material.h): every class lives in rtnw/scene.hpp
#pragma once
#include "rtnw/scene.hpp"
