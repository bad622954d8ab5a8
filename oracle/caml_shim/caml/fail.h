/* TEST INFRASTRUCTURE ONLY: stand-in for <caml/fail.h>. caml_failwith records the message and
 * longjmps to the wrapper, which reports it as an error code. */
#ifndef PPB_SHIM_FAIL_H
#define PPB_SHIM_FAIL_H
#include <setjmp.h>
extern __thread jmp_buf ppb_shim_fail_jmp;
extern __thread const char *ppb_shim_fail_msg;
static inline void caml_failwith(const char *msg) { ppb_shim_fail_msg = msg; longjmp(ppb_shim_fail_jmp, 1); }
#endif
