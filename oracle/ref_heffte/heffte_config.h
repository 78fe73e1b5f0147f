/* heffte_config.h -- TEST INFRASTRUCTURE ONLY (oracle/).  The vendored heFFTe
 * (/root/reference/heffte/heffteBenchmark) generates this header with CMake from
 * include/heffte_config.cmake.h; the reference's build system is not run here, so the three values a
 * CPU-only `stock`-backend build needs are written out by hand: version 2.1.0 (CMakeLists.txt project
 * version) and the AVX kernels of the stock backend (the library's fastest dependency-free CPU path). */
#ifndef HEFFTE_CONFIG_H
#define HEFFTE_CONFIG_H
#define Heffte_VERSION_MAJOR 2
#define Heffte_VERSION_MINOR 1
#define Heffte_VERSION_PATCH 0
#define Heffte_ENABLE_AVX
#endif
