// stand-in for the CMake-generated OSPConfig.h (modules/cpu/common/OSPConfig.h.in); default option values of
// modules/cpu/CMakeLists.txt:281-300
#pragma once
#define TILE_SIZE 64
#define MAX_TILE_SIZE 64
#define OSPRAY_RENDER_TASK_SIZE 8
