/* TEST INFRASTRUCTURE ONLY (oracle/). Minimal stand-in for OCaml's <caml/mlvalues.h>, just enough
 * to compile the reference's pplacer_src/linear_c.c UNMODIFIED into oracle/_ref/ so that the CPU
 * restatement in oracle/pplacer_oracle.c can be validated against the reference's own inner loops,
 * and to compile and drive integration/pplacer_b200_stubs.c (tests/test_ocaml_stub.py).
 * Written from the OCaml C-interface documentation; not OCaml runtime code. */
#ifndef PPB_SHIM_MLVALUES_H
#define PPB_SHIM_MLVALUES_H
#include <stdint.h>
typedef intptr_t value;
typedef intptr_t intnat;
#define CAMLprim
#define Val_unit ((value)1)
#define Int_val(v) ((int)((v) >> 1))
#define Val_int(i) ((value)(((intptr_t)(i) << 1) | 1))
/* a block is an (even) pointer to its fields; immediates are odd */
#define Is_block(v) (((v) & 1) == 0)
#define Field(v, i) (((value *)(v))[i])
#endif
