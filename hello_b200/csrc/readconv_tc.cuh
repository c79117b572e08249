// Fused tcgen05 read convolver (placeholder interface; implementation lands with the tensor-core path).
#pragma once
#include <string>
#include <vector>
#include "common.cuh"

namespace hello {

struct ReadConvTC;

static ReadConvTC* readconv_tc_create(const std::vector<LayerDesc>&, const float*, int, int, int, std::string& err) {
    err = "tensor-core read convolver not built in this revision; use HELLO_PREC_FP32";
    return nullptr;
}
static cudaError_t readconv_tc_launch(ReadConvTC*, const uint8_t*, long long, int, float*, cudaStream_t) {
    return cudaErrorNotSupported;
}
static void readconv_tc_destroy(ReadConvTC*) {}

}  // namespace hello
