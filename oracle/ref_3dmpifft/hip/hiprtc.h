/* hip/hiprtc.h -- TEST INFRASTRUCTURE ONLY: the run-time compiler the reference's FFT engine calls (templateFFT.cpp:5621-5705).
 * Here "compiling" a generated kernel = g++ -shared of its source against the HIP-on-CPU header (tfft_engine.cpp); the "code"
 * handed to hipModuleLoadDataEx is the path of that shared object. */
#ifndef REF3D_HIPRTC_SHIM_H
#define REF3D_HIPRTC_SHIM_H
#include <stddef.h>
typedef struct hiprtc_program_s* hiprtcProgram;
enum hiprtcResult { HIPRTC_SUCCESS = 0, HIPRTC_ERROR_COMPILATION = 6 };
#ifdef __cplusplus
extern "C" {
#endif
const char* hiprtcGetErrorString(enum hiprtcResult r);
enum hiprtcResult hiprtcCreateProgram(hiprtcProgram* prog, const char* src, const char* name, int nheaders, const char** headers, const char** names);
enum hiprtcResult hiprtcAddNameExpression(hiprtcProgram prog, const char* expr);
enum hiprtcResult hiprtcCompileProgram(hiprtcProgram prog, int nopt, const char** opts);
enum hiprtcResult hiprtcGetProgramLog(hiprtcProgram prog, char* log);
enum hiprtcResult hiprtcGetCodeSize(hiprtcProgram prog, size_t* n);
enum hiprtcResult hiprtcGetCode(hiprtcProgram prog, char* code);
enum hiprtcResult hiprtcDestroyProgram(hiprtcProgram* prog);
#ifdef __cplusplus
}
#endif
#endif
