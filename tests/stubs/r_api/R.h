/* R.h -- see Rinternals.h in this directory (declarations only, test infrastructure) */
#ifndef B200NN_STUB_R_H
#define B200NN_STUB_R_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#endif
