// Stand-in for <cufft.h> (see cuda_runtime.h in this directory): the emulated builds never execute a cuFFT plan - the
// fused derotation+FFT kernel or no FFT at all is used - so every call reports failure instead of pretending.
#pragma once
typedef int cufftHandle;
typedef int cufftResult;
typedef float2 cufftComplex;
enum { CUFFT_SUCCESS = 0, CUFFT_EMULATED = 1, CUFFT_C2C = 0x29, CUFFT_FORWARD = -1 };
static inline cufftResult cufftPlanMany(cufftHandle *, int, int *, int *, int, int, int *, int, int, int, int) { return CUFFT_EMULATED; }
static inline cufftResult cufftSetStream(cufftHandle, cudaStream_t) { return CUFFT_EMULATED; }
static inline cufftResult cufftExecC2C(cufftHandle, cufftComplex *, cufftComplex *, int) { return CUFFT_EMULATED; }
static inline cufftResult cufftDestroy(cufftHandle) { return CUFFT_SUCCESS; }
