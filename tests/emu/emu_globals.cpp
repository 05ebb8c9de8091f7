// Globals of the CPU emulation shim (tools/cuda_emu.h).  Test infrastructure only.
#include "../../tools/cuda_emu.h"
FH_EMU_DEFINE_GLOBALS

#if defined(__x86_64__)
// void fh_emu_switch(void** save_sp, void* load_sp): push the callee-saved registers, store the stack pointer, load the
// other stack, pop its registers and return into it (System V AMD64 ABI).
asm(R"(
.text
.globl fh_emu_switch
.type fh_emu_switch,@function
fh_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size fh_emu_switch,.-fh_emu_switch
)");
#endif
