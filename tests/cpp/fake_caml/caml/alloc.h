#pragma once
#include "mlvalues.h"
value caml_copy_int64(int64_t); value caml_copy_int32(int32_t); value caml_copy_nativeint(intnat); value caml_copy_double(double); value caml_copy_string(const char*);
value caml_alloc_tuple(int); value caml_alloc(int,int);
value caml_alloc_string(size_t);
