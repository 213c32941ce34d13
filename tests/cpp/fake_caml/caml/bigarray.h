#pragma once
#include "mlvalues.h"
struct caml_ba_array { void* data; intnat dim[1]; };
#define Caml_ba_array_val(v) ((struct caml_ba_array*)(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
#define CAML_BA_INT32 1
#define CAML_BA_C_LAYOUT 0
value caml_ba_alloc(int, int, void*, intnat*);
