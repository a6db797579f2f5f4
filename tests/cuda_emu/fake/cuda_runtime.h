// Found first on the include path when the kernel headers are compiled for the CPU emulator (tests/cuda_emu/cuda_emu.h);
// cuda_emu_rt.h adds host-side stand-ins for the runtime API (used only by the whole-library emulation, emu_lib).
#pragma once
#include "../cuda_emu.h"
#include "../cuda_emu_rt.h"
