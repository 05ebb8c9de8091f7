// Globals of the CPU emulation shim (tools/cuda_emu.h).  Test infrastructure only.
#include "../../tools/cuda_emu.h"
FH_EMU_DEFINE_GLOBALS
