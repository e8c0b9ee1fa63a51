// TEST-ONLY stand-in for MATLAB's mex.h: just enough declarations to syntax-check
// matlab/dmpc_b200_mex.cpp with g++ -fsyntax-only where no MATLAB exists.  Never linked.
#pragma once
#include <cstddef>
#include <cstdint>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef bool mxLogical;
enum mxClassID { mxDOUBLE_CLASS, mxINT32_CLASS };
enum mxComplexity { mxREAL };
bool mxIsChar(const mxArray*);
bool mxIsStruct(const mxArray*);
bool mxIsEmpty(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
mxArray* mxGetField(const mxArray*, mwSize, const char*);
double mxGetScalar(const mxArray*);
double* mxGetDoubles(const mxArray*);
void* mxGetData(const mxArray*);
mxLogical* mxGetLogicals(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
mxArray* mxCreateLogicalMatrix(mwSize, mwSize);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxDuplicateArray(const mxArray*);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
bool mexIsLocked();
void mexLock();
void mexUnlock();
int mexAtExit(void (*)(void));
