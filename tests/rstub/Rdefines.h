#include "Rinternals.h"
#define NEW_OBJECT(c) R_do_new_object(c)
#define MAKE_CLASS(n) R_do_MAKE_CLASS(n)
#define SET_SLOT(x, what, value) R_do_slot_assign(x, what, value)
#define GET_SLOT(x, what) R_do_slot(x, what)
