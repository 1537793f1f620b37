// x86-64 fiber switch for tests/host_emulation/simt.h: saves the callee-saved registers of the System V ABI on the current stack,
// stores the stack pointer, adopts the other fiber's stack and returns into it.  (MXCSR / x87 control words are not touched by
// the emulated code and are left alone.)
#if defined(__x86_64__)
asm(R"(
    .text
    .globl simt_switch
    .type simt_switch,@function
simt_switch:
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
    .size simt_switch, .-simt_switch
    .section .note.GNU-stack,"",@progbits
)");
#endif
