// internal helpers of libospcuda (not part of the C-ABI)
#pragma once
#include <cuda_runtime.h>
#include "../../../include/ospcuda.h"

int oc_set_error(cudaError_t e, const char *where);
static inline int oc_check(cudaError_t e, const char *where) { return e == cudaSuccess ? 0 : oc_set_error(e, where); }
void oc_count_launch();
#define OC_TRY(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return oc_set_error(_e, #expr);                                        \
  } while (0)
