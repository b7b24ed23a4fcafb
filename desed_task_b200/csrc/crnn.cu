// placeholder entry points - replaced as the CRNN kernels land
#include "common.cuh"
using namespace sedk;
extern "C" int sedk_crnn_forward(const sedk_crnn_plan*, void*) { SEDK_UNSUPPORTED("sedk_crnn_forward: not built yet"); }
extern "C" int sedk_crnn_backward(const sedk_crnn_plan*, void*) { SEDK_UNSUPPORTED("sedk_crnn_backward: not built yet"); }
extern "C" int sedk_sed_loss(const float*, const float*, const float*, const float*, const float*, const float*, int, int,
                             int, int, int, float, float*, float*, float*, void*) {
    SEDK_UNSUPPORTED("sedk_sed_loss: not built yet");
}
extern "C" int sedk_gemm(int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int,
                         const float*, int, void*) {
    SEDK_UNSUPPORTED("sedk_gemm: not built yet");
}
