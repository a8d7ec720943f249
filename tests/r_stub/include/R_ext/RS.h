/* see ../Rinternals.h: the one macro of R_ext/RS.h the reference's C sources use */
#ifndef RSTUB_RS_H
#define RSTUB_RS_H
#include <string.h>
#define Memzero(p, n) memset(p, 0, (size_t)(n) * sizeof(*p))
#endif
