// Quantized-conv forward (models/quantized_conv.py:36,38) -- placeholder until the kernels land.
#include "po2_common.cuh"

extern "C" {

size_t po2_conv2d_workspace(int, int, int, int, int, int, int, int, int, int, int) { return 0; }

int po2_conv2d_fwd(const void*, const void*, const float*, void*, int, int, int, int, int, int, int,
                   int, int, int, int, int, int, int, void*, size_t, void*) {
  return PO2_E_UNSUPPORTED;
}

}  // extern "C"
