/* see ../Rinternals.h: the registration types of R_ext/Rdynload.h */
#ifndef RSTUB_RDYNLOAD_H
#define RSTUB_RDYNLOAD_H
#include "../Rinternals.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef void *(*DL_FUNC)(void);
typedef struct {
    const char *name;
    DL_FUNC fun;
    int numArgs;
} R_CallMethodDef;
typedef struct rstub_dllinfo DllInfo;
int R_registerRoutines(DllInfo *info, const void *cRoutines, const R_CallMethodDef *callRoutines,
                       const void *fortranRoutines, const void *externalRoutines);
#ifdef __cplusplus
}
#endif
#endif
