/* shim: see mct_cvshim.h */
#include "mct_cvshim.h"
