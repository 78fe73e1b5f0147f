/* TEST INFRASTRUCTURE ONLY: the reference's FFT engine includes this header and uses nothing from it */
