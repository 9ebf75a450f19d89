/* Stub of MATLAB's mex.h, written for this repo (test infrastructure only).
 * It provides just the typedefs the reference's instruction table needs when it is
 * compiled with -DIGNORE_MEX (see /root/reference/lib/csparse/instructionsTableUTHash.c). */
#ifndef TC_ORACLE_STUB_MEX_H
#define TC_ORACLE_STUB_MEX_H
#include <stddef.h>
#include <stdio.h>
#include <stdbool.h>
typedef size_t mwSize;
typedef size_t mwIndex;
typedef bool boolean_T;
typedef struct mxArray_tag mxArray;
#define mexPrintf printf
#endif
