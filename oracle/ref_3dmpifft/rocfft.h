/* rocfft.h -- TEST INFRASTRUCTURE ONLY: fft_mpi_3d_api.cpp:338-377 still creates rocFFT plans that its execute path no longer
 * uses (the calls are commented out at :486-491, :526-533); plans are dummies, no work buffer is requested. */
#ifndef REF3D_ROCFFT_SHIM_H
#define REF3D_ROCFFT_SHIM_H
#include <stddef.h>
typedef struct rocfft_dummy_plan* rocfft_plan;
typedef struct rocfft_dummy_info* rocfft_execution_info;
typedef struct rocfft_dummy_desc* rocfft_plan_description;
typedef enum { rocfft_status_success = 0, rocfft_status_failure = 1 } rocfft_status;
typedef enum { rocfft_placement_inplace, rocfft_placement_notinplace } rocfft_result_placement;
typedef enum { rocfft_transform_type_complex_forward, rocfft_transform_type_complex_inverse } rocfft_transform_type;
typedef enum { rocfft_precision_single, rocfft_precision_double } rocfft_precision;
typedef enum { rocfft_array_type_complex_interleaved } rocfft_array_type;
static inline rocfft_status rocfft_plan_create(rocfft_plan* p, rocfft_result_placement, rocfft_transform_type, rocfft_precision, size_t, const size_t*, size_t, rocfft_plan_description) { *p = (rocfft_plan)(size_t)8; return rocfft_status_success; }
static inline rocfft_status rocfft_plan_destroy(rocfft_plan) { return rocfft_status_success; }
static inline rocfft_status rocfft_plan_get_work_buffer_size(rocfft_plan, size_t* n) { *n = 0; return rocfft_status_success; }
static inline rocfft_status rocfft_execution_info_create(rocfft_execution_info* i) { *i = (rocfft_execution_info)(size_t)8; return rocfft_status_success; }
static inline rocfft_status rocfft_execution_info_destroy(rocfft_execution_info) { return rocfft_status_success; }
static inline rocfft_status rocfft_execution_info_set_work_buffer(rocfft_execution_info, void*, size_t) { return rocfft_status_success; }
static inline rocfft_status rocfft_plan_description_create(rocfft_plan_description* d) { *d = (rocfft_plan_description)(size_t)8; return rocfft_status_success; }
static inline rocfft_status rocfft_plan_description_set_data_layout(rocfft_plan_description, rocfft_array_type, rocfft_array_type, const size_t*, const size_t*, size_t, const size_t*, size_t, size_t, const size_t*, size_t) { return rocfft_status_success; }
#endif
