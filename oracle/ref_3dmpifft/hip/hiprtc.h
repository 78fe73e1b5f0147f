/* hip/hiprtc.h -- TEST INFRASTRUCTURE ONLY: nothing of the run-time compiler is needed on the CPU (the FFT engine behind
 * templateFFT.h is replaced by ref3d_glue.cpp). */
