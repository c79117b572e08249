// Developer timeline of the generic tensor-core layer kernel (hello_b200/csrc/convlayer_tc.cuh): one layer of the 2x-wide read
// convolver on synthetic activations, CTA 0 stamps clock64() per tile and warp role.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/convlayer_trace.bin tools/convlayer_trace.cu
//   tools/convlayer_trace.bin [cin cout k stride pad n_items lin resid]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hello_b200/csrc/convlayer_tc.cuh"
using namespace hello;

int main(int argc, char** argv) {
    int cin = 64, cout = 64, k = 3, stride = 1, pad = 1, n_items = 38600, lin = 71, with_resid = 0;
    if (argc > 8) { cin = atoi(argv[1]); cout = atoi(argv[2]); k = atoi(argv[3]); stride = atoi(argv[4]); pad = atoi(argv[5]);
                    n_items = atoi(argv[6]); lin = atoi(argv[7]); with_resid = atoi(argv[8]); }
    std::string err;
    ConvLayerTC* t = convlayer_tc_create(HELLO_PREC_BF16X3, err);
    if (!t) { printf("create: %s\n", err.c_str()); return 1; }
    const size_t nw = (size_t)k * cin * cout;
    std::vector<float> hw(nw + cout);
    for (size_t i = 0; i < hw.size(); ++i) hw[i] = (float)((i * 2654435761u >> 8) % 2001) / 20000.f - 0.05f;
    float* dw; cudaMalloc(&dw, hw.size() * 4); cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    ConvDesc c{cin, cout, k, stride, pad, 1, dw, dw + nw};
    if (!convlayer_tc_add(t, c, dw, hw.data(), err, with_resid != 0) || !t->layers.count(dw)) { printf("add: %s (eligible %d)\n", err.c_str(), (int)cl::eligible(c)); return 1; }
    const int lout = c.out_len(lin);
    const size_t nx = (size_t)n_items * lin * cin, ny = (size_t)n_items * lout * cout;
    float *dx, *dy, *dr = nullptr;
    cudaMalloc(&dx, nx * 4); cudaMalloc(&dy, ny * 4);
    cudaMemset(dx, 0x3c, nx * 4);                         // 0x3c3c3c3c = 0.0115: finite activations
    if (with_resid) { cudaMalloc(&dr, ny * 4); cudaMemset(dr, 0x3c, ny * 4); }
    cudaMalloc(&t->d_trace, cl::TRACE_TILES * cl::TRACE_SLOTS * 8);
    ActView x = view_cl(dx, lin, cin);
    cudaError_t e;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(t->d_trace, 0, cl::TRACE_TILES * cl::TRACE_SLOTS * 8);
        cudaEventRecord(e0);
        if (!convlayer_tc_launch(t, x, c, n_items, dy, dr, 0, &e) || e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed\n"); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("rep %d: %.3f ms  (%.2f TB/s algorithmic)\n", rep, ms, ((nx + ny + (with_resid ? ny : 0)) * 4.0) / ms / 1e9);
    }
    std::vector<long long> tr(cl::TRACE_TILES * cl::TRACE_SLOTS);
    cudaMemcpy(tr.data(), t->d_trace, tr.size() * 8, cudaMemcpyDeviceToHost);
    const long long t0 = tr[0];
    printf("cycles since the first stamp; one row per tile of CTA 0\n");
    printf("%4s | %8s %8s %8s | %8s %8s %8s | %8s %8s || %7s %7s %7s %7s\n", "tile", "P:free", "P:loaded", "P:stored", "M:accfree", "M:opseen",
           "M:issued", "E:accseen", "E:done", "Pstore", "Mwait", "Missue", "Epi");
    for (int i = 0; i < cl::TRACE_TILES; ++i) {
        long long* s = &tr[i * cl::TRACE_SLOTS];
        if (!s[7]) break;
        printf("%4d | %8lld %8lld %8lld | %8lld %8lld %8lld | %8lld %8lld || %7lld %7lld %7lld %7lld\n", i, s[0] - t0, s[1] - t0, s[2] - t0,
               s[3] - t0, s[4] - t0, s[5] - t0, s[6] - t0, s[7] - t0, s[2] - s[0], s[4] - s[3], s[5] - s[4], s[7] - s[6]);
    }
    return 0;
}
