#pragma once
void caml_failwith(const char*); void caml_invalid_argument(const char*);
