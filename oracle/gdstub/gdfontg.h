/* oracle/gdstub/gdfontg.h -- TEST INFRASTRUCTURE ONLY: see gd.h in this directory. */
#include "gd.h"
