// std-backed Boost stand-in (test infrastructure), see _shim_core.hpp
#pragma once
#include "../_shim_core.hpp"
