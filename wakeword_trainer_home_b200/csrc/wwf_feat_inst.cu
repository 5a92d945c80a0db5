// wwf_feat_inst.cu - explicit instantiations of feat_kernel / feat_frames_kernel for ONE n_fft (-DWWF_INST_NFFT=...), so
// that build.py can compile the five n_fft families in parallel.  wwfeat.cu declares them extern.
#include "wwf_feat.cuh"

#ifndef WWF_INST_NFFT
#error "compile with -DWWF_INST_NFFT=<256|400|512|1024|2048>"
#endif

namespace wwf {
#define WWF_INST(H)                                                                      \
  template __global__ void feat_kernel<WWF_INST_NFFT, H, float>(const FeatParams);       \
  template __global__ void feat_kernel<WWF_INST_NFFT, H, __half>(const FeatParams);      \
  template __global__ void feat_frames_kernel<WWF_INST_NFFT, H>(const FeatParams);
WWF_INST(0)
WWF_INST(4)
WWF_INST(5)
#if WWF_INST_NFFT > 256
WWF_INST(8)
#endif
#if WWF_INST_NFFT >= 1024
WWF_INST(16)
#endif
}  // namespace wwf
