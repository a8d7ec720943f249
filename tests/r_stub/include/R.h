/* see Rinternals.h in this directory: a stand-in for R's headers used only by tests/test_r_shim.py */
#ifndef RSTUB_R_H
#define RSTUB_R_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#endif
