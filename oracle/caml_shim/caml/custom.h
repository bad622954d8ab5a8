/* TEST INFRASTRUCTURE ONLY: stand-in for <caml/custom.h> (nothing needed). */
#ifndef PPB_SHIM_CUSTOM_H
#define PPB_SHIM_CUSTOM_H
#endif
