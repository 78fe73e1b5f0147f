/* hipfft.h -- TEST INFRASTRUCTURE ONLY: fft_mpi_3d_api.cpp:318-336 still creates (and :150-157 destroys) hipFFT plans that
 * its execute path no longer uses; the handles are dummies. */
#ifndef REF3D_HIPFFT_SHIM_H
#define REF3D_HIPFFT_SHIM_H
typedef struct hipfft_dummy_plan* hipfftHandle;
typedef int hipfftResult_t;
typedef hipfftResult_t hipfftResult;
#define HIPFFT_SUCCESS 0
#define HIPFFT_Z2Z 0x69
#define HIPFFT_FORWARD (-1)
#define HIPFFT_BACKWARD 1
typedef struct { double x, y; } hipfftDoubleComplex;
static inline hipfftResult_t hipfftPlanMany(hipfftHandle* p, int, int*, int*, int, int, int*, int, int, int, int) { *p = (hipfftHandle)(size_t)8; return 0; }
static inline hipfftResult_t hipfftPlan2d(hipfftHandle* p, int, int, int) { *p = (hipfftHandle)(size_t)8; return 0; }
static inline hipfftResult_t hipfftDestroy(hipfftHandle) { return 0; }
#endif
