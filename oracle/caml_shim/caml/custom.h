/* TEST INFRASTRUCTURE ONLY: stand-in for <caml/custom.h>: a custom block is {ops pointer, payload...}. */
#ifndef PPB_SHIM_CUSTOM_H
#define PPB_SHIM_CUSTOM_H
#include <stdlib.h>
#include "mlvalues.h"
struct custom_operations {
  const char *identifier;
  void (*finalize)(value);
  void *compare, *hash, *serialize, *deserialize, *compare_ext, *fixed_length;
};
#define custom_compare_default NULL
#define custom_hash_default NULL
#define custom_serialize_default NULL
#define custom_deserialize_default NULL
#define custom_compare_ext_default NULL
#define custom_fixed_length_default NULL
#define Data_custom_val(v) ((void *)((value *)(v) + 1))
static inline value caml_alloc_custom(struct custom_operations *ops, unsigned long size, unsigned long mem, unsigned long max) {
  value *b = (value *)calloc(1, sizeof(value) + size);
  (void)mem; (void)max;
  b[0] = (value)ops;
  return (value)b;
}
/* what the GC would do when the block dies: run the finaliser and free (called by the test driver) */
static inline void ppb_shim_finalize_custom(value v) {
  struct custom_operations *ops = (struct custom_operations *)((value *)v)[0];
  if (ops && ops->finalize) ops->finalize(v);
  free((void *)v);
}
#endif
