/* R_ext/Rdynload.h -- see ../Rinternals.h (declarations only, test infrastructure) */
#ifndef B200NN_STUB_RDYNLOAD_H
#define B200NN_STUB_RDYNLOAD_H
typedef void *(*DL_FUNC)(void);
typedef struct {
  const char *name;
  DL_FUNC fun;
  int numArgs;
} R_CallMethodDef;
typedef struct _DllInfo DllInfo;
typedef R_CallMethodDef R_CMethodDef, R_FortranMethodDef, R_ExternalMethodDef;
int R_registerRoutines(DllInfo *, const R_CMethodDef *, const R_CallMethodDef *, const R_FortranMethodDef *,
                       const R_ExternalMethodDef *);
int R_useDynamicSymbols(DllInfo *, int);
#endif
