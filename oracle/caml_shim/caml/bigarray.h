/* TEST INFRASTRUCTURE ONLY: stand-in for <caml/bigarray.h>. A "bigarray value" in this shim is a
 * pointer to a shim_ba descriptor {data, dim[]} built by oracle/ref_wrap.c. */
#ifndef PPB_SHIM_BIGARRAY_H
#define PPB_SHIM_BIGARRAY_H
#include "mlvalues.h"
struct caml_ba_array { void *data; intnat dim[4]; };
#define Caml_ba_array_val(v) ((struct caml_ba_array *)(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
static inline intnat caml_ba_num_elts(struct caml_ba_array *b) {
  intnat n = 1; int i;
  for (i = 0; i < 4 && b->dim[i] > 0; ++i) n *= b->dim[i];
  return n;
}
#endif
