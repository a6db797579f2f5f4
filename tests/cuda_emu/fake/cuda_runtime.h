// Found first on the include path when the kernel headers are compiled for the CPU emulator (tests/cuda_emu/cuda_emu.h).
#pragma once
#include "../cuda_emu.h"
