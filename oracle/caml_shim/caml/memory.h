/* TEST INFRASTRUCTURE ONLY: stand-in for <caml/memory.h>; no GC here so the root macros are no-ops. */
#ifndef PPB_SHIM_MEMORY_H
#define PPB_SHIM_MEMORY_H
#include "mlvalues.h"
#define CAMLparam1(a) (void)(a)
#define CAMLparam2(a,b) (void)(a);(void)(b)
#define CAMLparam3(a,b,c) (void)(a);(void)(b);(void)(c)
#define CAMLparam4(a,b,c,d) (void)(a);(void)(b);(void)(c);(void)(d)
#define CAMLparam5(a,b,c,d,e) (void)(a);(void)(b);(void)(c);(void)(d);(void)(e)
#define CAMLxparam1(a) (void)(a)
#define CAMLxparam2(a,b) (void)(a);(void)(b)
#define CAMLxparam3(a,b,c) (void)(a);(void)(b);(void)(c)
#define CAMLxparam4(a,b,c,d) (void)(a);(void)(b);(void)(c);(void)(d)
#define CAMLxparam5(a,b,c,d,e) (void)(a);(void)(b);(void)(c);(void)(d);(void)(e)
#define CAMLlocal1(x) value x = 0
#define CAMLreturn(x) return (x)
#endif
