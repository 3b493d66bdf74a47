// fyn_conv_tc.cu -- tcgen05 / TMEM implicit-GEMM convolution family (placeholder: not yet enabled).
#include "fyn_internal.h"

int fyn_conv_tc_supported(const fyn_conv_desc *, int) { return 0; }
int fyn_conv_tc_create(fyn_op *, const float *) { FYN_FAIL(FYN_ERR_UNSUPPORTED, "tcgen05 family not built"); }
int fyn_conv_tc_run(fyn_op *, const fyn_tensor *, const fyn_tensor *, fyn_tensor *, cudaStream_t) { return 1; }
void fyn_conv_tc_destroy(fyn_op *) {}
