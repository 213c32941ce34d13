/* STAND-IN for the OCaml runtime headers (the image has no OCaml): just enough declarations for gcc -fsyntax-only
   to check ocaml/coherence_stubs.c against include/coherence_b200.h.  Test infrastructure only. */
#pragma once
#include <stdint.h>
#include <stddef.h>
typedef intptr_t value; typedef intptr_t intnat;
#define CAMLprim
#define Val_unit ((value)1)
#define Val_int(x) ((value)(x))
#define Val_long(x) ((value)(x))
#define Int_val(v) ((int)(v))
#define Long_val(v) ((long)(v))
#define Bool_val(v) ((int)(v))
#define Val_bool(x) ((value)((x) != 0))
#define Field(v,i) (((value*)(v))[i])
#define Store_field(b,i,v) (((value*)(b))[i]=(v))
#define Wosize_val(v) ((size_t)((value*)(v))[-1])
#define Double_wosize 1
#define Double_val(v) (*(double*)(v))
#define Double_flat_field(v,i) (((double*)(v))[i])
#define String_val(v) ((const char*)(v))
#define Nativeint_val(v) (*(intnat*)(v))
#define Int64_val(v) (*(int64_t*)(v))
#define Int32_val(v) (*(int32_t*)(v))
#define Bytes_val(v) ((unsigned char*)(v))
size_t caml_string_length(value);
