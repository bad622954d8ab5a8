/* TEST INFRASTRUCTURE ONLY: stand-in for <caml/alloc.h>. caml_copy_double "boxes" into a
 * thread-local slot; oracle/ref_wrap.c unboxes it immediately. */
#ifndef PPB_SHIM_ALLOC_H
#define PPB_SHIM_ALLOC_H
#include "mlvalues.h"
static __thread double ppb_shim_boxed_double;
static inline value caml_copy_double(double d) { ppb_shim_boxed_double = d; return (value)&ppb_shim_boxed_double; }
#define Double_val(v) (*(double *)(v))
#endif
