// Fused tcgen05 combiner: ConcatenateChannels -> Conv1d(256 -> 512, k=3, pad=1) + ReLU -> Conv1d(512 -> 128, k=1) + ReLU
// (reference: python/architectures/conv_combiner.py:10-42, used by MoEAttention.forward at allele level (combiner0) and
// site level (combiner1), python/MixtureOfExpertsAdvanced.py:193-219) in ONE persistent kernel.  Two fp32 [n,18,128]
// tensors go in (the channel concat is just "which tensor a K-half comes from"), fp32 [n,18,128] comes out.
//
// Same operand scheme as headconv_tc.cuh (8-channel chunk arrays of 16-byte rows, taps = row shifts, 6 items packed
// with pitch 20 into one 128-row tile, weights streamed from L2 through a 6 x 16 KB ring in units of one tap x 16
// input channels x 128 outputs).  What is particular here is the size: the 512-channel intermediate of 128 rows is
// 256 KB as bf16 hi+lo and cannot sit in shared memory, and the 256-channel input takes 125 KB.  So
//   * the first convolution is K-split: the operand buffer holds one 128-channel half (tensor a, then tensor b) and
//     all 512 accumulator columns of TMEM collect both halves;
//   * the intermediate is consumed in four chunks of 128 channels: the epilogue turns accumulator columns
//     [128q, 128q+128) into the operand T_q (64 KB) and the second convolution accumulates T_q x W2[128q.., :] into
//     D2, which reuses the accumulator columns [0,128) the first chunk has already vacated.
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hello_moe.h"
#include "common.cuh"
#include "tc_ptx.cuh"
#include "readconv_tc.cuh"   // store_chunk8, bf16 host helpers, HostConv
#include "headconv_tc.cuh"   // ring constants

namespace hello {
namespace cc {

constexpr int NSLOT = hc::NSLOT;
constexpr uint32_t SLOT_BYTES = hc::SLOT_BYTES;
constexpr int C_HALF = 128, C_MID = 512, C_OUT = 128, L = 18, P = 20, G = 6, ROWS = G * P;
constexpr int EPI_WARPS = 8, THREADS = (EPI_WARPS + 2) * 32;
constexpr uint32_t X_ARR = (ROWS + 2) * 16;               // first-conv operand: one array per 8-channel chunk, lead row
constexpr uint32_t X_LO = (C_HALF / 8) * X_ARR;
constexpr uint32_t T_ARR = 128 * 16;                      // second-conv operand (1x1: no halo rows)
constexpr uint32_t T_LO = (128 / 8) * T_ARR;
constexpr uint32_t OFF_X = 0;
constexpr uint32_t OFF_T = 2 * X_LO;
constexpr uint32_t OFF_W = OFF_T + 2 * T_LO;
constexpr uint32_t OFF_BIAS = OFF_W + NSLOT * SLOT_BYTES;
constexpr int N_BIAS = C_MID + C_OUT;
constexpr uint32_t OFF_BAR = OFF_BIAS + N_BIAS * 4;
constexpr int BAR_FULL = 0, BAR_EMPTY = NSLOT, BAR_X = 2 * NSLOT, BAR_HALF = BAR_X + 1, BAR_ACC = BAR_X + 2,
              BAR_T = BAR_X + 3, BAR_PB = BAR_X + 4, N_BARS = BAR_X + 5;
constexpr uint32_t OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16;
static_assert(OFF_T % 128 == 0 && OFF_W % 128 == 0 && SMEM_BYTES <= 232448, "shared memory budget");
constexpr int UNITS_A = 3 * (C_HALF / 16);                // units of one (K-half, N-chunk) of the first conv
constexpr int UNITS_B = 128 / 16;                         // units of one K-chunk of the second conv
constexpr int N_UNITS = 2 * 4 * UNITS_A + 4 * UNITS_B;    // per work item, in consumption order

struct CombParams {
    const float* in_a;        // K-half 0: [n, 18, .] fp32, `in_stride` floats between rows
    const float* in_b;        // K-half 1
    const uint8_t* weights;   // N_UNITS packed units in consumption order
    const float* bias;        // [512 + 128]
    float* out;               // [n, 18, 128]
    float* dbg;               // optional [groups][128][512] dump of the intermediate (dbg_phase 0)
    long long n_items;
    int in_stride;
    int n_work;
    int dbg_phase;
};

template <int MODE> __host__ __device__ constexpr uint32_t unit_bytes() { return (MODE == 3 ? 2u : 1u) * 128u * 32u; }

// fp32 rows of one 128-channel half of the group -> first-conv operand (bf16 hi + lo chunk arrays, lead row kept zero)
template <int MODE>
__device__ __forceinline__ void load_half(uint8_t* xbuf, const float* __restrict__ src, int stride, long long i0, int n, int tid) {
    constexpr int CH8 = C_HALF / 8, TOTAL = ROWS * CH8, NT = EPI_WARPS * 32;
    for (int base = tid; base < TOTAL; base += NT * 4) {
        float4 va[4][2];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * NT;
            const int c8 = idx % CH8, m = idx / CH8;
            const int i = m / P, p = m - i * P;
            ok[u] = idx < TOTAL && i < n && p < L;
            if (ok[u]) {
                const float4* s4 = reinterpret_cast<const float4*>(src + ((i0 + i) * L + p) * (long long)stride + c8 * 8);
                va[u][0] = __ldg(s4); va[u][1] = __ldg(s4 + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * NT;
            if (idx >= TOTAL) continue;
            const int c8 = idx % CH8, m = idx / CH8;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ok[u]) {
                v[0] = va[u][0].x; v[1] = va[u][0].y; v[2] = va[u][0].z; v[3] = va[u][0].w;
                v[4] = va[u][1].x; v[5] = va[u][1].y; v[6] = va[u][1].z; v[7] = va[u][1].w;
            }
            tc::store_chunk8<MODE>(xbuf + c8 * X_ARR + (uint32_t)(m + 1) * 16, X_LO, v);
        }
    }
}

// `n_units` weight units from the ring against the A operand at act_lo; unit u reads tap u / k16 (row shift) and
// channel step u % k16.  Accumulates into d (first unit overwrites when `fresh`).
template <int MODE>
__device__ __forceinline__ void issue_units(int n_units, int k16, uint32_t act_lo, uint32_t a_lbo, uint32_t a_lo_plane,
                                            uint32_t d, bool fresh, uint32_t ring_lo, uint32_t bar_full0,
                                            uint32_t bar_empty0, uint32_t& slot, uint32_t& par) {
    constexpr uint32_t UNIT = unit_bytes<MODE>(), UNIT_HI = 128u * 32u;
    constexpr int UPF = SLOT_BYTES / UNIT;
    constexpr uint32_t idesc = ptx::idesc_bf16_m128(128);
#pragma unroll 1
    for (int u0 = 0; u0 < n_units; u0 += UPF) {
        ptx::mbar_wait(bar_full0 + 8u * slot, par);
#pragma unroll
        for (int k = 0; k < UPF; ++k) {
            const int u = u0 + k;
            const int tap = u / k16, j = u - tap * k16;
            const uint32_t a = (act_lo + (uint32_t)tap + (uint32_t)j * ((2 * a_lbo) >> 4)) | (((a_lbo >> 4) & 0x3FFFu) << 16);
            const uint32_t b = (ring_lo + ((slot * SLOT_BYTES + (uint32_t)k * UNIT) >> 4)) | (((128u * 16u) >> 4) << 16);
            const uint32_t acc = (fresh && u == 0) ? 0u : 1u;
            if (MODE == 3) {
                ptx::mma_bf16_ss(d, a + (a_lo_plane >> 4), b, idesc, acc);
                ptx::mma_bf16_ss(d, a, b + (UNIT_HI >> 4), idesc, 1u);
                ptx::mma_bf16_ss(d, a, b, idesc, 1u);
            } else {
                ptx::mma_bf16_ss(d, a, b, idesc, acc);
            }
        }
        ptx::tc_commit(bar_empty0 + 8u * slot);
        __syncwarp();
        if (++slot == NSLOT) { slot = 0; par ^= 1u; }
    }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) combconv_tc_kernel(const __grid_constant__ CombParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
    const uint32_t bar0 = ptx::smem_u32(smem + OFF_BAR);
    auto bar = [&](int k) { return bar0 + 8u * k; };
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { ptx::mbar_init(bar(BAR_FULL + s), 1); ptx::mbar_init(bar(BAR_EMPTY + s), 1); }
        ptx::mbar_init(bar(BAR_X), EPI_WARPS * 32);
        ptx::mbar_init(bar(BAR_HALF), 1);
        ptx::mbar_init(bar(BAR_ACC), 1);
        ptx::mbar_init(bar(BAR_T), EPI_WARPS * 32);
        ptx::mbar_init(bar(BAR_PB), 1);
        ptx::fence_mbar_init();
    }
    for (int i = threadIdx.x; i < N_BIAS; i += blockDim.x) s_bias[i] = __ldg(prm.bias + i);
    {
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (uint32_t i = threadIdx.x; i < OFF_BIAS / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (warp == EPI_WARPS + 1) {
        ptx::tmem_alloc(ptx::smem_u32(smem + OFF_TMEM), 512);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
    const long long n_items = prm.n_items;

    if (warp < EPI_WARPS) {
        // ===================================================== loader / epilogue warps
        const int wq = warp & 3, chalf = warp >> 2, wrow = wq * 32, tid = threadIdx.x;
        uint8_t* xbuf = smem + OFF_X;
        uint8_t* tbuf = smem + OFF_T;
        const uint32_t tl = tmem_base + ((uint32_t)wrow << 16);
        const int m = wrow + lane;
        const int mi = m / P, mp = m - mi * P;
        uint32_t half_n = 0, acc_n = 0, pb_n = 0;
        for (int item = blockIdx.x; item < prm.n_work; item += gridDim.x) {
            const long long i0 = (long long)item * G;
            const int n = (int)min((long long)G, n_items - i0);
            const bool valid = mi < n && mp < L;
            load_half<MODE>(xbuf, prm.in_a, prm.in_stride, i0, n, tid);
            ptx::tc_fence_before();
            ptx::fence_proxy_async();
            ptx::mbar_arrive(bar(BAR_X));
            ptx::mbar_wait(bar(BAR_HALF), half_n & 1u);            // MMAs of the first K-half have read the operand
            ++half_n;
            load_half<MODE>(xbuf, prm.in_b, prm.in_stride, i0, n, tid);
            ptx::tc_fence_before();
            ptx::fence_proxy_async();
            ptx::mbar_arrive(bar(BAR_X));
            ptx::mbar_wait(bar(BAR_ACC), acc_n & 1u);              // all 512 accumulator columns complete
            ++acc_n;
            ptx::tc_fence_after();
            float* dbg = (prm.dbg && prm.dbg_phase == 0) ? prm.dbg + (long long)item * (128 * C_MID) : nullptr;
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                if (q > 0) {                                       // T is free once the previous chunk's MMAs are done
                    ptx::mbar_wait(bar(BAR_PB), pb_n & 1u);
                    ++pb_n;
                    ptx::tc_fence_after();
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    const int c0 = chalf * 64 + b * 32;
                    float v[32];
                    ptx::tmem_ld32(tl + q * 128 + c0, v);
                    ptx::tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float x = fmaxf(v[c] + s_bias[q * 128 + c0 + c], 0.f);
                        v[c] = valid ? x : 0.f;
                    }
                    if (dbg) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) dbg[m * C_MID + q * 128 + c0 + c] = v[c];
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::store_chunk8<MODE>(tbuf + (c0 / 8 + k) * T_ARR + (uint32_t)m * 16, T_LO, v + 8 * k);
                }
                ptx::tc_fence_before();
                ptx::fence_proxy_async();
                ptx::mbar_arrive(bar(BAR_T));
            }
            ptx::mbar_wait(bar(BAR_PB), pb_n & 1u);                // D2 complete
            ++pb_n;
            ptx::tc_fence_after();
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int c0 = chalf * 64 + b * 32;
                float v[32];
                ptx::tmem_ld32(tl + c0, v);
                ptx::tmem_wait_ld();
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(prm.out + ((i0 + mi) * L + mp) * (long long)C_OUT + c0);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        dst[k] = make_float4(fmaxf(v[4 * k] + s_bias[C_MID + c0 + 4 * k], 0.f),
                                             fmaxf(v[4 * k + 1] + s_bias[C_MID + c0 + 4 * k + 1], 0.f),
                                             fmaxf(v[4 * k + 2] + s_bias[C_MID + c0 + 4 * k + 2], 0.f),
                                             fmaxf(v[4 * k + 3] + s_bias[C_MID + c0 + 4 * k + 3], 0.f));
                }
            }
        }
    } else if (warp == EPI_WARPS) {
        // ===================================================== MMA issuer
        uint32_t slot = 0, par = 0, x_n = 0, t_n = 0;
        const uint32_t x_lo = ptx::smem_u32(smem + OFF_X) >> 4;
        const uint32_t t_lo = ptx::smem_u32(smem + OFF_T) >> 4;
        const uint32_t ring_lo = ptx::smem_u32(smem + OFF_W) >> 4;
        for (int item = blockIdx.x; item < prm.n_work; item += gridDim.x) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                ptx::mbar_wait(bar(BAR_X), x_n & 1u);
                ++x_n;
                ptx::tc_fence_after();
#pragma unroll 1
                for (int q = 0; q < 4; ++q)
                    issue_units<MODE>(UNITS_A, C_HALF / 16, x_lo, X_ARR, X_LO, tmem_base + q * 128, h == 0, ring_lo,
                                      bar(BAR_FULL), bar(BAR_EMPTY), slot, par);
                ptx::tc_commit(bar(h == 0 ? BAR_HALF : BAR_ACC));
                __syncwarp();
            }
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                ptx::mbar_wait(bar(BAR_T), t_n & 1u);
                ++t_n;
                ptx::tc_fence_after();
                issue_units<MODE>(UNITS_B, UNITS_B, t_lo, T_ARR, T_LO, tmem_base, q == 0, ring_lo, bar(BAR_FULL),
                                  bar(BAR_EMPTY), slot, par);
                ptx::tc_commit(bar(BAR_PB));
                __syncwarp();
            }
        }
    } else {
        // ===================================================== weight producer
        if (lane == 0) {
            uint32_t slot = 0, par = 1;
            const uint32_t w0 = ptx::smem_u32(smem + OFF_W);
            const uint32_t total = N_UNITS * unit_bytes<MODE>();
            for (int item = blockIdx.x; item < prm.n_work; item += gridDim.x) {
                for (uint32_t o = 0; o < total; o += SLOT_BYTES) {
                    ptx::mbar_wait(bar(BAR_EMPTY + slot), par);
                    ptx::mbar_expect_tx(bar(BAR_FULL + slot), SLOT_BYTES);
                    ptx::bulk_g2s(w0 + slot * SLOT_BYTES, prm.weights + o, 8192u, bar(BAR_FULL + slot));
                    ptx::bulk_g2s(w0 + slot * SLOT_BYTES + 8192u, prm.weights + o + 8192u, 8192u, bar(BAR_FULL + slot));
                    if (++slot == NSLOT) { slot = 0; par ^= 1u; }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + 1) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace cc

struct CombConvTC {
    cc::CombParams prm;
    uint8_t* d_weights = nullptr;
    float* d_bias = nullptr;
    int mode = 3, sm_count = 148;
};

// Checks that `net` is "Conv(256 -> 512, k3, pad 1, ReLU), Conv(512 -> 128, k1, ReLU)" and packs its weights into ring
// units in the order the kernel consumes them.
static CombConvTC* combconv_tc_create(const std::vector<LayerDesc>& net, const float* d_base, const float* h_base,
                                      int precision, std::string& err) {
    using namespace cc;
    if (precision != HELLO_PREC_BF16X3 && precision != HELLO_PREC_BF16) { err = "unknown tensor-core precision"; return nullptr; }
    auto is_conv = [&](const LayerDesc& Ld, int cin, int cout, int k, int pad) {
        return Ld.kind == KIND_CONV && Ld.a.cin == cin && Ld.a.cout == cout && Ld.a.k == k && Ld.a.stride == 1 && Ld.a.pad == pad && Ld.a.relu == ACT_RELU;
    };
    if (net.size() != 2 || !is_conv(net[0], 2 * C_HALF, C_MID, 3, 1) || !is_conv(net[1], C_MID, C_OUT, 1, 0)) {
        err = "layer table is not the 256 -> 512 (k3) -> 128 (k1) combiner"; return nullptr;
    }
    const int parts = precision == HELLO_PREC_BF16X3 ? 2 : 1;
    auto hcv = [&](const ConvDesc& c) { return tc::HostConv{h_base + (c.w - d_base), h_base + (c.b - d_base), c.cin, c.cout, c.k}; };
    const tc::HostConv c1 = hcv(net[0].a), c2 = hcv(net[1].a);
    std::vector<uint16_t> blob;
    // unit: [hi: 2 chunks][128 rows n][8] then lo; element (chunk c, n, e) = W[n0 + n][ci0 + 8c + e][tap]
    auto pack_unit = [&](const tc::HostConv& c, int tap, int ci0, int n0) {
        std::vector<uint16_t> hi, lo;
        for (int ch = 0; ch < 2; ++ch)
            for (int n = 0; n < 128; ++n)
                for (int e = 0; e < 8; ++e) {
                    const float w = c.w[(size_t)(tap * c.cin + ci0 + 8 * ch + e) * c.cout + n0 + n];
                    const uint16_t h = tc::bf16_rne(w);
                    hi.push_back(h);
                    lo.push_back(tc::bf16_rne(w - tc::bf16_to_float(h)));
                }
        blob.insert(blob.end(), hi.begin(), hi.end());
        if (parts == 2) blob.insert(blob.end(), lo.begin(), lo.end());
    };
    for (int h = 0; h < 2; ++h)
        for (int q = 0; q < 4; ++q)
            for (int tap = 0; tap < 3; ++tap)
                for (int j = 0; j < C_HALF / 16; ++j) pack_unit(c1, tap, h * C_HALF + 16 * j, q * 128);
    for (int q = 0; q < 4; ++q)
        for (int j = 0; j < UNITS_B; ++j) pack_unit(c2, 0, q * 128 + 16 * j, 0);
    if (blob.size() * 2 != (size_t)N_UNITS * 128 * 32 * parts) { err = "combiner weight packing size mismatch"; return nullptr; }
    std::vector<float> bias(N_BIAS);
    for (int i = 0; i < C_MID; ++i) bias[i] = c1.b[i];
    for (int i = 0; i < C_OUT; ++i) bias[C_MID + i] = c2.b[i];

    CombConvTC* t = new CombConvTC();
    t->mode = parts == 2 ? 3 : 1;
    std::memset(&t->prm, 0, sizeof(t->prm));
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { err = "cudaGetDeviceProperties failed"; delete t; return nullptr; }
    t->sm_count = prop.multiProcessorCount;
    if ((size_t)prop.sharedMemPerBlockOptin < SMEM_BYTES) { err = "device has too little shared memory per block"; delete t; return nullptr; }
    if (cudaMalloc(&t->d_weights, blob.size() * 2) != cudaSuccess || cudaMalloc(&t->d_bias, bias.size() * 4) != cudaSuccess ||
        cudaMemcpy(t->d_weights, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t->d_bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        err = "allocating the packed bf16 combiner weights failed";
        if (t->d_weights) cudaFree(t->d_weights);
        if (t->d_bias) cudaFree(t->d_bias);
        delete t;
        return nullptr;
    }
    cudaError_t e = t->mode == 3
        ? cudaFuncSetAttribute(combconv_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)
        : cudaFuncSetAttribute(combconv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) {
        err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        cudaFree(t->d_weights); cudaFree(t->d_bias); delete t;
        return nullptr;
    }
    t->prm.weights = t->d_weights;
    t->prm.bias = t->d_bias;
    return t;
}

// in_a / in_b: the two 128-channel halves, fp32, `in_stride` floats between consecutive rows (128 for two separate
// [n,18,128] tensors, 256 for the halves of one concatenated [n,18,256] tensor).  out: fp32 [n, 18, 128].
static cudaError_t combconv_tc_launch(CombConvTC* t, const float* in_a, const float* in_b, int in_stride, long long n,
                                      float* out, cudaStream_t st, float* dbg = nullptr, int dbg_phase = -1) {
    if (n <= 0) return cudaSuccess;
    cc::CombParams prm = t->prm;
    prm.in_a = in_a; prm.in_b = in_b; prm.in_stride = in_stride; prm.n_items = n; prm.out = out;
    prm.dbg = dbg; prm.dbg_phase = dbg_phase;
    const long long work = (n + cc::G - 1) / cc::G;
    if (work > 0x7fffffffLL) return cudaErrorInvalidValue;
    prm.n_work = (int)work;
    const int grid = (int)std::min<long long>(work, t->sm_count);
    if (t->mode == 3) cc::combconv_tc_kernel<3><<<grid, cc::THREADS, cc::SMEM_BYTES, st>>>(prm);
    else cc::combconv_tc_kernel<1><<<grid, cc::THREADS, cc::SMEM_BYTES, st>>>(prm);
    return cudaGetLastError();
}

static void combconv_tc_destroy(CombConvTC* t) {
    if (!t) return;
    if (t->d_weights) cudaFree(t->d_weights);
    if (t->d_bias) cudaFree(t->d_bias);
    delete t;
}

}  // namespace hello
