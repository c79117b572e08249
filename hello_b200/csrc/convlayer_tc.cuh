// Generic tcgen05 Conv1d layer (+bias, +ReLU, +residual): the tensor-core counterpart of conv_fp32.cuh for every layer
// that is not inside one of the fused kernels -- meta_convolver_ref (architectures/meta_convolver_ref.py:13-106) and
// all sub-networks of the 2x-wide models (architectures/*_wide.py), which do not fit the fused kernels' shared-memory
// layouts.  Replaces torch.nn.Conv1d called through NNTools.WeightNormedConv1d (python/NNTools.py:791-799).
//
// Implicit GEMM, one 128-row x NT-column output tile per CTA (rows = (item, output position), NT = min(Cout, 256)),
// K = taps x Cin walked in steps of 16 channels through a 4-stage ring:
//   warps 0-3  A producers: thread = output row; gathers the 16 fp32 input channels of its row for the step's tap
//              (zero outside the item: the padding), splits them into bf16 hi + lo and writes the K-major operand
//              tile; after the last step the same warps are the epilogue (TMEM -> bias / ReLU / residual -> HBM)
//   warp 4     one thread streams the step's packed weight unit with cp.async.bulk
//   warp 5     MMA issuer (3 products per step in bf16x3 mode), frees a ring stage with tcgen05.commit
// Two CTAs fit an SM (<= 97 KB shared memory, <= 256 TMEM columns each), so one CTA's epilogue overlaps the other's
// MMAs.  Layer by layer through HBM: these layers are a few percent of a step; the hot stacks have fused kernels.
#pragma once
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hello_moe.h"
#include "common.cuh"
#include "tc_ptx.cuh"
#include "readconv_tc.cuh"   // bf16 host helpers

namespace hello {
namespace cl {

constexpr int STAGES = 4, THREADS = 192;
constexpr uint32_t A_STAGE = 8192;                       // 128 rows x 16 channels x (hi, lo) bf16

struct ConvTcArgs {
    const float* x;          // fp32 channel-last input [n_items][lin][cin] (item stride sn, row stride sl floats)
    long long sn, sl;
    const uint8_t* w;        // packed units: [n tile][tap][k16]( [hi: 2 chunks][nt][8], [lo: ...] )
    const float* bias;
    float* y;                // [M][cout]
    const float* resid;      // [M][cout] or nullptr, added after the ReLU
    long long M;
    int lin, lout, cin, cout, ksz, stride, pad, relu, nt, tmem_cols;
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) convlayer_tc_kernel(const ConvTcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t unit_hi = 32u * a.nt, unit = MODE == 3 ? 2 * unit_hi : unit_hi;
    const uint32_t a_stage = MODE == 3 ? A_STAGE : A_STAGE / 2;
    uint8_t* s_a = smem;
    uint8_t* s_b = smem + STAGES * a_stage;
    const uint32_t bar0 = ptx::smem_u32(s_b + STAGES * unit);
    auto bar = [&](int k) { return bar0 + 8u * k; };
    constexpr int BAR_AFULL = 0, BAR_BFULL = STAGES, BAR_EMPTY = 2 * STAGES, BAR_ACC = 3 * STAGES;
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(s_b + STAGES * unit + (3 * STAGES + 1) * 8);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(bar(BAR_AFULL + s), 128); ptx::mbar_init(bar(BAR_BFULL + s), 1); ptx::mbar_init(bar(BAR_EMPTY + s), 1);
        }
        ptx::mbar_init(bar(BAR_ACC), 1);
        ptx::fence_mbar_init();
    }
    if (warp == 5) {
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(s_tmem)), (uint32_t)a.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const int k16 = a.cin / 16, steps = a.ksz * k16;
    const long long m0 = (long long)blockIdx.x * 128;
    const int n0 = blockIdx.y * a.nt;

    if (warp < 4) {
        // ------------------------------------------------------------------ A producers, then epilogue
        const int r = threadIdx.x;
        const long long m = m0 + r;
        const bool row_ok = m < a.M;
        long long item = 0;
        int pos0 = 0;
        if (row_ok) { item = m / a.lout; pos0 = (int)(m - item * a.lout) * a.stride - a.pad; }
        const float* xitem = a.x + item * a.sn;
        uint32_t stage = 0, par = 1;
        // The gather of step s + PF is issued as soon as step s has been stored: PF - 1 steps of MMA time cover the
        // latency of the global loads (one step is only ~200 tensor-pipe cycles).
        constexpr int PF = 3;
        float v[PF][16];
        auto gather = [&](int s, float (&dst)[16]) {
            const int tap = s / k16, j = s - tap * k16;
            const int pos = pos0 + tap;
            if (row_ok && pos >= 0 && pos < a.lin) {
                const float4* p = reinterpret_cast<const float4*>(xitem + (long long)pos * a.sl + 16 * j);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = __ldg(p + q);
                    dst[4 * q] = t.x; dst[4 * q + 1] = t.y; dst[4 * q + 2] = t.z; dst[4 * q + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) dst[q] = 0.f;
            }
        };
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if (u < steps) gather(u, v[u]);
        for (int s0 = 0; s0 < steps; s0 += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                const int s = s0 + u;
                if (s < steps) {
                    ptx::mbar_wait(bar(BAR_EMPTY + stage), par);
                    uint8_t* dst = s_a + stage * a_stage + (uint32_t)r * 16;
                    tc::store_chunk8<MODE>(dst, A_STAGE / 2, v[u]);              // channels 0-7: chunk 0 (hi; lo plane 4 KB further)
                    tc::store_chunk8<MODE>(dst + 2048, A_STAGE / 2, v[u] + 8);   // channels 8-15: chunk 1
                    ptx::fence_proxy_async();
                    ptx::mbar_arrive(bar(BAR_AFULL + stage));
                    if (s + PF < steps) gather(s + PF, v[u]);
                    if (++stage == STAGES) { stage = 0; par ^= 1u; }
                }
            }
        }
        ptx::mbar_wait(bar(BAR_ACC), 0);
        ptx::tc_fence_after();
        const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16);
        float* yrow = a.y + m * a.cout + n0;
        const float* rrow = a.resid ? a.resid + m * a.cout + n0 : nullptr;
        for (int c0 = 0; c0 < a.nt; c0 += 16) {
            float acc[16];
            ptx::tmem_ld16(tl + c0, acc);
            ptx::tmem_wait_ld();
            if (row_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + c0) + q);
                    float4 o = make_float4(acc[4 * q] + b.x, acc[4 * q + 1] + b.y, acc[4 * q + 2] + b.z, acc[4 * q + 3] + b.w);
                    if (a.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    if (rrow) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(rrow + c0) + q);
                        o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                    }
                    reinterpret_cast<float4*>(yrow + c0)[q] = o;
                }
            }
        }
    } else if (warp == 4) {
        // ------------------------------------------------------------------ weight producer
        if (lane == 0) {
            uint32_t stage = 0, par = 1;
            const uint8_t* src = a.w + (size_t)blockIdx.y * steps * unit;
            for (int s = 0; s < steps; ++s) {
                ptx::mbar_wait(bar(BAR_EMPTY + stage), par);
                ptx::mbar_expect_tx(bar(BAR_BFULL + stage), unit);
                for (uint32_t o = 0; o < unit; o += 8192u)
                    ptx::bulk_g2s(ptx::smem_u32(s_b + stage * unit) + o, src + (size_t)s * unit + o, min(8192u, unit - o),
                                  bar(BAR_BFULL + stage));
                if (++stage == STAGES) { stage = 0; par ^= 1u; }
            }
        }
    } else {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = ptx::idesc_bf16_m128((uint32_t)a.nt);
        uint32_t stage = 0, par = 0;
        for (int s = 0; s < steps; ++s) {
            ptx::mbar_wait(bar(BAR_AFULL + stage), par);
            ptx::mbar_wait(bar(BAR_BFULL + stage), par);
            ptx::tc_fence_after();
            const uint32_t al = ptx::desc_lo(ptx::smem_u32(s_a + stage * a_stage), 2048);
            const uint32_t bl = ptx::desc_lo(ptx::smem_u32(s_b + stage * unit), (uint32_t)a.nt * 16u);
            if (MODE == 3) {
                ptx::mma_bf16_ss(tmem, al + ((A_STAGE / 2) >> 4), bl, idesc, s == 0 ? 0u : 1u);   // lo * hi
                ptx::mma_bf16_ss(tmem, al, bl + (unit_hi >> 4), idesc, 1u);                         // hi * lo
                ptx::mma_bf16_ss(tmem, al, bl, idesc, 1u);                                          // hi * hi
            } else {
                ptx::mma_bf16_ss(tmem, al, bl, idesc, s == 0 ? 0u : 1u);
            }
            ptx::tc_commit(bar(BAR_EMPTY + stage));
            __syncwarp();
            if (++stage == STAGES) { stage = 0; par ^= 1u; }
        }
        ptx::tc_commit(bar(BAR_ACC));
        __syncwarp();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 5) ptx::tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

struct PackedConv {
    uint8_t* d_w = nullptr;
    int nt = 0, tmem_cols = 0;
};

inline bool eligible(const ConvDesc& c) {
    return c.cin % 16 == 0 && c.cout % 16 == 0 && c.cin >= 16 && c.cout >= 16 && (c.k == 1 || c.k == 3) &&
           (c.cout <= 256 || c.cout % 256 == 0);
}

}  // namespace cl

// Packed weights of every generic tensor-core layer of a handle, keyed by the layer's (device) weight pointer.
struct ConvLayerTC {
    std::map<const float*, cl::PackedConv> layers;
    int mode = 3;
};

static bool convlayer_tc_add(ConvLayerTC* t, const ConvDesc& c, const float* d_base, const float* h_base, std::string& err) {
    if (!cl::eligible(c) || t->layers.count(c.w)) return true;
    const int parts = t->mode == 3 ? 2 : 1;
    const float* w = h_base + (c.w - d_base);                       // [k*cin][cout]
    const int nt = c.cout <= 256 ? c.cout : 256;
    std::vector<uint16_t> blob;
    for (int n0 = 0; n0 < c.cout; n0 += nt)
        for (int tap = 0; tap < c.k; ++tap)
            for (int j = 0; j < c.cin / 16; ++j)
                for (int part = 0; part < parts; ++part)
                    for (int ch = 0; ch < 2; ++ch)
                        for (int n = 0; n < nt; ++n)
                            for (int e = 0; e < 8; ++e) {
                                const float v = w[(size_t)(tap * c.cin + 16 * j + 8 * ch + e) * c.cout + n0 + n];
                                const uint16_t h = tc::bf16_rne(v);
                                blob.push_back(part == 0 ? h : tc::bf16_rne(v - tc::bf16_to_float(h)));
                            }
    cl::PackedConv p;
    p.nt = nt;
    p.tmem_cols = nt <= 32 ? 32 : nt <= 64 ? 64 : nt <= 128 ? 128 : 256;
    if (cudaMalloc(&p.d_w, blob.size() * 2) != cudaSuccess ||
        cudaMemcpy(p.d_w, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
        err = "allocating packed layer weights failed";
        if (p.d_w) cudaFree(p.d_w);
        return false;
    }
    t->layers[c.w] = p;
    return true;
}

static size_t convlayer_tc_smem(int mode, int nt) {
    const size_t unit = (mode == 3 ? 64u : 32u) * nt, a_stage = mode == 3 ? cl::A_STAGE : cl::A_STAGE / 2;
    return cl::STAGES * (a_stage + unit) + (3 * cl::STAGES + 1) * 8 + 16;
}

static ConvLayerTC* convlayer_tc_create(int precision, std::string& err) {
    ConvLayerTC* t = new ConvLayerTC();
    t->mode = precision == HELLO_PREC_BF16X3 ? 3 : 1;
    const int smem = (int)convlayer_tc_smem(t->mode, 256);
    cudaError_t e = t->mode == 3
        ? cudaFuncSetAttribute(cl::convlayer_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
        : cudaFuncSetAttribute(cl::convlayer_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); delete t; return nullptr; }
    return t;
}

// true when the layer ran on tensor cores; false = not covered here (caller falls back to the fp32 kernel)
static bool convlayer_tc_launch(ConvLayerTC* t, const ActView& x, const ConvDesc& c, long long n_items, float* y,
                                const float* resid, cudaStream_t st, cudaError_t* e) {
    *e = cudaSuccess;
    if (!t || x.is_u8 || x.sc != 1 || x.ch != c.cin || (x.sl % 4) || (x.sn % 4) ||
        (reinterpret_cast<uintptr_t>(x.base) & 15)) return false;
    auto it = t->layers.find(c.w);
    if (it == t->layers.end()) return false;
    if (n_items <= 0) return true;
    cl::ConvTcArgs a;
    a.x = static_cast<const float*>(x.base); a.sn = x.sn; a.sl = x.sl;
    a.w = it->second.d_w; a.bias = c.b; a.y = y; a.resid = resid;
    a.lin = x.len; a.lout = c.out_len(x.len); a.M = n_items * a.lout;
    a.cin = c.cin; a.cout = c.cout; a.ksz = c.k; a.stride = c.stride; a.pad = c.pad; a.relu = c.relu;
    a.nt = it->second.nt; a.tmem_cols = it->second.tmem_cols;
    const long long tiles = (a.M + 127) / 128;
    if (tiles > 0x7fffffffLL) { *e = cudaErrorInvalidValue; return true; }
    dim3 grid((unsigned)tiles, (unsigned)(c.cout / a.nt));
    const size_t smem = convlayer_tc_smem(t->mode, a.nt);
    if (t->mode == 3) cl::convlayer_tc_kernel<3><<<grid, cl::THREADS, smem, st>>>(a);
    else cl::convlayer_tc_kernel<1><<<grid, cl::THREADS, smem, st>>>(a);
    *e = cudaGetLastError();
    return true;
}

static void convlayer_tc_destroy(ConvLayerTC* t) {
    if (!t) return;
    for (auto& kv : t->layers) if (kv.second.d_w) cudaFree(kv.second.d_w);
    delete t;
}

}  // namespace hello
