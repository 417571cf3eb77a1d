// aus_kernels.cu -- the kernels of ONE fft_size (2^AUS_LG), both precisions, every variant, with their launch functions.
// Compiled once per size (Makefile: -DAUS_LG=6 .. 14) so the nine sizes build in parallel.
#include "aus_host.cuh"

#ifndef AUS_LG
#error "compile with -DAUS_LG=<log2 fft_size, 6..14>"
#endif

template int launch_stft_t<double, AUS_LG>(aus_ctx *, const StftParams<double> &);
template int launch_stft_t<float, AUS_LG>(aus_ctx *, const StftParams<float> &);
template int launch_analyzer_t<double, AUS_LG>(aus_ctx *, const StftParams<double> &);
template int launch_analyzer_t<float, AUS_LG>(aus_ctx *, const StftParams<float> &);
template int launch_istft_t<double, AUS_LG>(aus_ctx *, const IstftParams<double> &);
template int launch_istft_t<float, AUS_LG>(aus_ctx *, const IstftParams<float> &);
