// Shared declarations for the hello_moe CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "hello_moe is written for sm_100a (B200) only"
#endif

namespace hello {

// Activation after a convolution: the `relu` field of a layer record (the architecture modules' `activation` switch,
// python/NNTools.py:72-115).  Softplus = torch.nn.Softplus() with its defaults (beta 1, threshold 20).
enum Activation { ACT_NONE = 0, ACT_RELU = 1, ACT_SOFTPLUS = 2 };

__host__ __device__ __forceinline__ float apply_activation(float v, int act) {
#ifdef __CUDA_ARCH__
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_SOFTPLUS) return v > 20.f ? v : log1pf(expf(v));
#endif
    return v;
}

// Softplus inside the fused tensor-core kernels, where the epilogue is on the critical path: two special-function
// instructions (ex2.approx, lg2.approx) instead of the ~40 of log1pf(expf(v)); absolute error < 2e-7, an order of
// magnitude under the bf16 hi+lo split of the value it feeds.
__device__ __forceinline__ float softplus_fast(float v) {
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));     // exp(v); flushes to 0 below -87
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + e));                      // the argument is >= 1: never subnormal
    return v > 20.f ? v : l * 0.6931471805599453f;
}

// One convolution (or linear) layer, weights already weight-norm folded.
struct ConvDesc {
    int cin, cout, k, stride, pad, relu;   // relu: an Activation code
    const float* w;     // device: conv [k*cin][cout] (row = tap*cin + ci); linear [cout][cin]
    const float* b;     // device: [cout]
    int out_len(int lin) const { return (lin + 2 * pad - k) / stride + 1; }
};

enum LayerKind { KIND_CONV = 0, KIND_MAXPOOL = 1, KIND_RES = 2, KIND_GAP_LINEAR = 3 };

struct LayerDesc {
    int kind;
    int has_shortcut;
    ConvDesc a, b, s;   // conv: a; maxpool: a.k/a.stride; res: a, b, (s); gap_linear: a
};

// Where the activations of a layer's input live: element (item n, position p, channel c) is at
// base[n*sn + p*sl + c*sc]; `is_u8` selects uint8 vs fp32 elements.
struct ActView {
    const void* base;
    long long sn, sl, sc;
    int len, ch;
    bool is_u8;
};

static inline ActView view_cl(const float* p, int len, int ch) {  // channel-last fp32 [n][len][ch]
    ActView v;
    v.base = p; v.sn = (long long)len * ch; v.sl = ch; v.sc = 1; v.len = len; v.ch = ch; v.is_u8 = false;
    return v;
}

}  // namespace hello
