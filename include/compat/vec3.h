// Forwarding header: the reference's "vec3.h" is provided by include/rt_host.hpp (add -Iinclude -Iinclude/compat).
#pragma once
#include "../rt_host.hpp"
