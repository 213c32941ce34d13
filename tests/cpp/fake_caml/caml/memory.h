#pragma once
#define CAMLparam0()
#define CAMLparam1(a)
#define CAMLparam2(a,b)
#define CAMLparam3(a,b,c)
#define CAMLparam4(a,b,c,d)
#define CAMLparam5(a,b,c,d,e)
#define CAMLxparam1(a)
#define CAMLxparam2(a,b)
#define CAMLlocal1(a) value a = 0
#define CAMLlocal2(a,b) value a = 0, b = 0
#define CAMLlocal3(a,b,c) value a = 0, b = 0, c = 0
#define CAMLreturn(x) return (x)
