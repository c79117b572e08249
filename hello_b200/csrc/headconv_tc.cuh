// Fused tcgen05 allele/site head: one persistent kernel runs a whole "1x1 conv -> stride-2 residual block ->
// 2 identity residual blocks (-> mean over positions -> linear)" sub-network of HELLO's MoE forward:
//   C = 64   compressor{0,1}    python/architectures/compressor_conv_small.py:8-55   [A,36,64]  -> [A,18,128]
//   C = 128  xattn{0,1,2}       python/architectures/xattn_subtract.py:9-95          [A,18,128] -> [A,1]
//            (front end 2*allele - site, LinearCombination[2,-1], folded into the operand loader)
//   C = 128  meta_convolver     python/architectures/meta_convolver.py:10-77         [S,18,128] -> softmax [S,3]
// fp32 features go in, fp32 features / logits come out; the seven layer phases in between never touch HBM.
//
// Same operand scheme as the read convolver (readconv_tc.cuh): activations sit in shared memory as 8-channel
// chunk arrays of 16-byte rows (K-major, no swizzle), a k=3 tap is the same array read one row further down, the
// stride-2 block reads an even/odd de-interleaved copy written by the previous epilogue, items are packed back to
// back along M (pitch 40 -> 20 or 20 -> 10 rows, the spare rows stay zero and are the pad=1 halo), accumulators and the
// fp32 residual stream live in TMEM, operands are rewritten in place by the epilogue.
// What differs: the layers are wide (N = 128 / 256, K up to 768), so one layer's weights (up to 786 KB as bf16
// hi+lo) do not fit next to the activations.  A producer thread streams them from L2 through a 6 x 16 KB ring
// with cp.async.bulk in "units" (one tap x 16 input channels x all N outputs, hi then lo); the MMA issuer consumes
// unit by unit (3 MMAs of 128 x N x 16 each in bf16x3 mode).  The MMAs of these layers are 64-128 tensor-pipe
// cycles each, so issue cost is irrelevant and the issuer is a plain loop.
//
//   C = 64 : 2 groups x 6 items in flight (TMEM 2 x 256 columns), 4 epilogue warps per group
//   C = 128: 1 group x 12 items (TMEM 512 columns: 256 accumulator + 256 residual), 8 epilogue warps, two per TMEM
//            lane quadrant, each owning half of the output columns
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hello_moe.h"
#include "common.cuh"
#include "tc_ptx.cuh"
#include "readconv_tc.cuh"   // store_chunk8, bf16 host helpers, HostConv

namespace hello {
namespace hc {

constexpr int NSLOT = 6;
constexpr uint32_t SLOT_BYTES = 16384;
constexpr int N_PHASES = 7;                    // layer phases of the standard head
// Addendum models (architectures/compressor_conv_small_addendum.py, xattn_subtract_addendum.py) append residual blocks at
// 2C channels in front of the output / pooled linear head: two more phases each, the instructions of phases 3..6.
constexpr int MAX_EXTRA_BLOCKS = 2;
constexpr int MAX_PHASES = N_PHASES + 2 * MAX_EXTRA_BLOCKS;
constexpr int TRACE_SLOTS = MAX_PHASES + 1;    // timeline record: one slot per phase + one for the operand loader
constexpr int N_BIAS_MAX = (15 + 4 * MAX_EXTRA_BLOCKS) * 128;
constexpr int EPI_WARPS = 8;
constexpr int DBG_ROWS = 256, DBG_COLS = 256;

template <int C>
struct Geo {
    static constexpr int L = C == 64 ? 36 : 18;       // input positions per item
    static constexpr int L2 = L / 2;                  // after the stride-2 block
    static constexpr int G = C == 64 ? 6 : 12;        // items per group
    static constexpr int NGRP = C == 64 ? 2 : 1;      // groups in flight per CTA
    static constexpr int CS = 2 / NGRP;               // epilogue warps per TMEM lane quadrant (column split)
    static constexpr int P1 = C == 64 ? 40 : 20, P2 = P1 / 2;   // row pitch of one item at the two resolutions
    static constexpr int ROWS1 = G * P1, ROWS2 = G * P2;
    static constexpr int N2 = 2 * C;
    static constexpr uint32_t X_ARR = 241 * 16;               // phase-0 operand: one array per 8-channel chunk
    static constexpr uint32_t E_ARR = (ROWS2 + 2) * 16;       // phase-0 output: (chunk, parity) arrays, lead row
    static constexpr uint32_t S_ARR = (ROWS2 + 2) * 16;       // stage-2 operands: one array per chunk, lead row
    static constexpr uint32_t X_LO = (C / 8) * X_ARR;         // distance to the "lo" plane of each layout
    static constexpr uint32_t E_LO = (C / 8) * 2 * E_ARR;
    static constexpr uint32_t S_LO = (N2 / 8) * S_ARR;
    static constexpr uint32_t ACT_BYTES = 2 * (E_LO > S_LO ? (E_LO > X_LO ? E_LO : X_LO) : (S_LO > X_LO ? S_LO : X_LO));
    // bias table (floats): 1x1 conv, block-1 conv a / shortcut / conv b, then two convs per following block
    static constexpr int B_0 = 0, B_1A = C, B_1S = 3 * C, B_1B = 5 * C, B_REST = 7 * C;
    static constexpr uint32_t OFF_ACT = 0;
    static constexpr uint32_t OFF_W = NGRP * ACT_BYTES;
    static constexpr uint32_t OFF_BAR = OFF_W + NSLOT * SLOT_BYTES;
    static constexpr uint32_t N_BARS = 2 * NSLOT + 2 * NGRP;  // full[NSLOT], empty[NSLOT], act_ready[NGRP], acc_full[NGRP]
    static constexpr uint32_t OFF_TMEM = OFF_BAR + N_BARS * 8;
    static constexpr uint32_t SMEM_BYTES = OFF_TMEM + 16;
    static constexpr int THREADS = (EPI_WARPS + NGRP + 1) * 32;
    static_assert(ROWS1 <= 256 && ROWS2 <= 128 && ROWS1 <= 241, "tiles cover the packed rows");
    static_assert(ACT_BYTES % 128 == 0 && SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(NGRP * 4 * C <= 512, "TMEM budget");
};

struct HeadParams {
    const float* in_a;        // [n, L, C] fp32 channel-last
    const float* in_s;        // optional [n_sites, L, C]: operand = 2*in_a[i] - in_s[site_idx[i]]
    const int32_t* site_idx;  // [n] (with in_s)
    const uint8_t* weights;   // packed units, phase after phase
    float bias_tab[N_BIAS_MAX];   // biases travel in the kernel parameters (constant bank), like the read convolver's
    const float* lin_w;       // [n_out][2C] pooled linear head (n_out > 0)
    const float* lin_b;       // [n_out]
    float* out;               // n_out == 0: [n, L/2, 2C];  else out[i*out_stride + o]
    float* dbg;               // optional [groups][256][256] dump of phase dbg_phase
    long long n_items;
    long long out_stride;
    uint32_t w_src[MAX_PHASES];   // byte offset of each phase's first unit
    int n_phases;             // 7 + 2 per appended residual block
    int n_work;               // work items (NGRP groups each)
    int n_out, softmax, dbg_phase;
};

enum { OUT_NAT = 0, OUT_EO = 1, OUT_FINAL = 2 };

// fp32 rows of one group -> phase-0 operand (bf16 hi + lo chunk arrays); invalid rows are written as zeros
template <int MODE, int C>
__device__ __forceinline__ void load_input(uint8_t* act, const HeadParams& prm, long long i0, int n, int tid) {
    using Gm = Geo<C>;
    constexpr int CH8 = C / 8, TOTAL = Gm::ROWS1 * CH8, NT = 128 * Gm::CS, UNR = 5;
    static_assert(TOTAL % (NT * UNR) == 0, "loader tiling");
    const bool front = prm.in_s != nullptr;
    for (int base = tid; base < TOTAL; base += NT * UNR) {
        float4 va[UNR][2], vs[UNR][2];
        bool ok[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int idx = base + u * NT;
            const int c8 = idx % CH8, m = idx / CH8;
            const int i = m / Gm::P1, p = m - i * Gm::P1;
            ok[u] = i < n && p < Gm::L;
            if (ok[u]) {
                const float4* src = reinterpret_cast<const float4*>(prm.in_a + ((i0 + i) * Gm::L + p) * (long long)C + c8 * 8);
                va[u][0] = __ldg(src); va[u][1] = __ldg(src + 1);
                if (front) {
                    const long long s = __ldg(prm.site_idx + i0 + i);
                    const float4* ss = reinterpret_cast<const float4*>(prm.in_s + (s * Gm::L + p) * (long long)C + c8 * 8);
                    vs[u][0] = __ldg(ss); vs[u][1] = __ldg(ss + 1);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int idx = base + u * NT;
            const int c8 = idx % CH8, m = idx / CH8;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ok[u]) {
                v[0] = va[u][0].x; v[1] = va[u][0].y; v[2] = va[u][0].z; v[3] = va[u][0].w;
                v[4] = va[u][1].x; v[5] = va[u][1].y; v[6] = va[u][1].z; v[7] = va[u][1].w;
                if (front) {   // LinearCombination[2,-1]: one rounding of 2a - s
                    const float s[8] = {vs[u][0].x, vs[u][0].y, vs[u][0].z, vs[u][0].w,
                                        vs[u][1].x, vs[u][1].y, vs[u][1].z, vs[u][1].w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaf(2.f, v[e], -s[e]);
                }
            }
            tc::store_chunk8<MODE>(act + c8 * Gm::X_ARR + (uint32_t)m * 16, Gm::X_LO, v);
        }
    }
}

// Epilogue of one convolution for one group.  Thread = TMEM lane = packed row; this warp owns columns
// [chalf*N/CS, (chalf+1)*N/CS).   y = relu(acc + bias) [+ resid (+ bias2)], invalid rows forced to zero.
template <int MODE, int C, int ACT, int N, int TILES, int PITCH, int LVALID, bool RESID, bool RES_BIAS, bool WRITE_RESID, int OUT>
__device__ __forceinline__ void epi_conv(uint8_t* act, uint32_t tl, int bias, int bias2, int n,
                                         uint32_t out_stride, uint32_t out_lo, const HeadParams& prm, long long i0,
                                         float* __restrict__ dbg, int wrow, int lane, int chalf, float* part) {
    using Gm = Geo<C>;
    constexpr int NB = N / Gm::CS / 32;          // 32-column blocks per warp
    constexpr int ROWS = Gm::G * PITCH;
    constexpr uint32_t RES_COL = Gm::N2;
    for (int tile = 0; tile < TILES; ++tile) {
        const int m = tile * 128 + wrow + lane;
        const int i = m / PITCH, p = m - i * PITCH;
        const bool valid = (i < n) && (p < LVALID);
        const bool in_buf = m < ROWS;
#pragma unroll 1                                  // blocks are serial anyway (TMEM load wait); keeps the kernel ~2x smaller
        for (int b = 0; b < NB; ++b) {
            const int c0 = (chalf * NB + b) * 32;
            float v[32];
            float r[RESID ? 32 : 1];
            ptx::tmem_ld32(tl + tile * N + c0, v);
            if (RESID) ptx::tmem_ld32(tl + RES_COL + c0, r);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float x = v[c] + prm.bias_tab[bias + c0 + c];
                x = ACT == ACT_RELU ? fmaxf(x, 0.f) : softplus_fast(x);
                if (RESID) x += RES_BIAS ? (r[c] + prm.bias_tab[bias2 + c0 + c]) : r[c];
                v[c] = valid ? x : 0.f;
            }
            if (WRITE_RESID) ptx::tmem_st32(tl + RES_COL + c0, v);
            if (dbg) {
#pragma unroll
                for (int c = 0; c < 32; ++c) dbg[m * DBG_COLS + c0 + c] = v[c];
            }
            if (OUT == OUT_FINAL) {
                if (prm.n_out == 0) {
                    if (valid) {
                        float4* dst = reinterpret_cast<float4*>(prm.out + ((i0 + i) * Gm::L2 + p) * (long long)N + c0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                } else {
                    for (int o = 0; o < prm.n_out; ++o) {
                        const float4* w4 = reinterpret_cast<const float4*>(prm.lin_w + o * N + c0);
                        float s = 0.f;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 w = __ldg(w4 + q);
                            s = fmaf(v[4 * q], w.x, s); s = fmaf(v[4 * q + 1], w.y, s);
                            s = fmaf(v[4 * q + 2], w.z, s); s = fmaf(v[4 * q + 3], w.w, s);
                        }
                        part[o] += s;
                    }
                }
            } else if (in_buf) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c8 = c0 / 8 + q;
                    uint8_t* dst = OUT == OUT_NAT ? act + c8 * out_stride + (uint32_t)(m + 1) * 16
                                                  : act + (c8 * 2 + (m & 1)) * out_stride + (uint32_t)((m >> 1) + 1) * 16;
                    tc::store_chunk8<MODE>(dst, out_lo, v + 8 * q);
                }
            }
        }
    }
    if (OUT != OUT_FINAL && wrow + lane == 0) {          // zero padding row in front of the first item
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        constexpr int ARRS = (N / Gm::CS / 8) * (OUT == OUT_EO ? 2 : 1);
        const int a0 = chalf * ARRS;
#pragma unroll 4
        for (int a = a0; a < a0 + ARRS; ++a) {
            *reinterpret_cast<uint4*>(act + a * out_stride) = z;
            if (MODE == 3) *reinterpret_cast<uint4*>(act + a * out_stride + out_lo) = z;
        }
    }
    if (WRITE_RESID) ptx::tmem_wait_st();
}

// All tcgen05.mma of one layer phase for one group, unit by unit as the weights arrive in the ring.
//   KIND 0: 1x1 conv C -> C on the two stage-1 tiles;  KIND 1: stride-2 conv a (3 taps) + 1x1 stride-2 shortcut
//   (accumulates straight into the residual columns);  KIND 2: k=3 conv 2C -> 2C.
template <int MODE, int C, int KIND>
__device__ __forceinline__ void issue_phase(bool active, uint32_t act_lo, uint32_t ring_lo, uint32_t d_acc,
                                            uint32_t bar_full0, uint32_t bar_empty0, uint32_t& slot, uint32_t& par,
                                            int lane) {
    using Gm = Geo<C>;
    constexpr int N = KIND == 0 ? C : 2 * C;
    constexpr int K16 = KIND == 2 ? (2 * C) / 16 : C / 16;
    constexpr int UNITS = KIND == 0 ? K16 : KIND == 1 ? 4 * K16 : 3 * K16;
    constexpr uint32_t UNIT_HI = 32u * N, UNIT_BYTES = MODE == 3 ? 2 * UNIT_HI : UNIT_HI;
    constexpr int UPF = SLOT_BYTES / UNIT_BYTES;
    constexpr int TILES = KIND == 0 ? 2 : 1;
    constexpr uint32_t idesc = ptx::idesc_bf16_m128(N);
    constexpr uint32_t A_LBO = KIND == 0 ? Gm::X_ARR : KIND == 1 ? 2 * Gm::E_ARR : Gm::S_ARR;
    constexpr uint32_t A_LO = KIND == 0 ? Gm::X_LO : KIND == 1 ? Gm::E_LO : Gm::S_LO;
    const uint32_t d_res = d_acc + Gm::N2;
#pragma unroll 1
    for (int u0 = 0; u0 < UNITS; u0 += UPF) {
        ptx::mbar_wait(bar_full0 + 8u * slot, par);
        if (active) {
#pragma unroll
            for (int k = 0; k < UPF; ++k) {
                const int u = u0 + k;
                if (u < UNITS) {
                    uint32_t a_off, d;
                    bool first;
                    if (KIND == 0) {
                        a_off = (uint32_t)u * (2 * A_LBO); d = d_acc; first = u == 0;
                    } else if (KIND == 1) {
                        const int tap = u / K16, j = u - tap * K16;
                        // stride 2: x[2q-1], x[2q], x[2q+1] = odd[q-1], even[q], odd[q]; the shortcut reads even[q]
                        const uint32_t ao = tap == 0 ? Gm::E_ARR : tap == 2 ? Gm::E_ARR + 16u : 16u;
                        a_off = ao + (uint32_t)j * (2 * A_LBO);
                        d = tap < 3 ? d_acc : d_res;
                        first = tap < 3 ? u == 0 : j == 0;
                    } else {
                        const int tap = u / K16, j = u - tap * K16;
                        a_off = (uint32_t)tap * 16u + (uint32_t)j * (2 * A_LBO); d = d_acc; first = u == 0;
                    }
                    const uint32_t a0 = (act_lo + (a_off >> 4)) | (((A_LBO >> 4) & 0x3FFFu) << 16);
                    const uint32_t b0 = (ring_lo + ((slot * SLOT_BYTES + (uint32_t)k * UNIT_BYTES) >> 4)) |
                                        ((((uint32_t)N * 16u) >> 4) << 16);
#pragma unroll
                    for (int t = 0; t < TILES; ++t) {
                        const uint32_t at = a0 + (uint32_t)t * 128u, dt = d + (uint32_t)t * C;
                        if (MODE == 3) {                        // small terms first: lo*hi, hi*lo, then hi*hi
                            ptx::mma_bf16_ss(dt, at + (A_LO >> 4), b0, idesc, first ? 0u : 1u);
                            ptx::mma_bf16_ss(dt, at, b0 + (UNIT_HI >> 4), idesc, 1u);
                            ptx::mma_bf16_ss(dt, at, b0, idesc, 1u);
                        } else {
                            ptx::mma_bf16_ss(dt, at, b0, idesc, first ? 0u : 1u);
                        }
                    }
                }
            }
            ptx::tc_commit(bar_empty0 + 8u * slot);          // ring slot no longer read by this group
        } else if (lane == 0) {
            ptx::mbar_arrive(bar_empty0 + 8u * slot);
        }
        __syncwarp();
        if (++slot == NSLOT) { slot = 0; par ^= 1u; }
    }
}

template <int MODE, int C>
__device__ __forceinline__ uint32_t phase_bytes(int ph) {
    const uint32_t per16 = (MODE == 3 ? 64u : 32u);       // bytes per (unit, output channel)
    const uint32_t k16 = C / 16;
    return ph == 0 ? k16 * per16 * C : ph == 1 ? 4 * k16 * per16 * 2 * C : 6 * k16 * per16 * 2 * C;
}

// Timeline hook (dbg_phase == -2, debug instantiation only): CTA 0 stamps clock64() for its first 16 work items as int64
// [item][group][TRACE_SLOTS][4] = {issue start, issue end, accumulators seen, epilogue end}; slot MAX_PHASES holds
// {operand load start, operand load end, 0, 0}.
__device__ __forceinline__ long long* head_trace_slot(const HeadParams& prm, int item, int g, int ngrp) {
    if (!prm.dbg || prm.dbg_phase != -2 || blockIdx.x != 0) return nullptr;
    const int li = item / (int)gridDim.x;
    if (li >= 16) return nullptr;
    return reinterpret_cast<long long*>(prm.dbg) + ((long long)(li * ngrp + g) * TRACE_SLOTS) * 4;
}

template <int MODE, int C, bool DBG, int ACT = ACT_RELU>
__global__ void __launch_bounds__(Geo<C>::THREADS, 1) headconv_tc_kernel(const __grid_constant__ HeadParams prm) {
    using Gm = Geo<C>;
    constexpr int NGRP = Gm::NGRP, CS = Gm::CS, G = Gm::G;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t bar0 = ptx::smem_u32(smem + Gm::OFF_BAR);
    auto bar = [&](int k) { return bar0 + 8u * k; };
    constexpr int BAR_FULL = 0, BAR_EMPTY = NSLOT, BAR_ACT = 2 * NSLOT, BAR_ACC = 2 * NSLOT + NGRP;
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + Gm::OFF_TMEM);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { ptx::mbar_init(bar(BAR_FULL + s), 1); ptx::mbar_init(bar(BAR_EMPTY + s), NGRP); }
        for (int g = 0; g < NGRP; ++g) { ptx::mbar_init(bar(BAR_ACT + g), 128 * CS); ptx::mbar_init(bar(BAR_ACC + g), 1); }
        ptx::fence_mbar_init();
    }
    {   // rows the MMAs read past the written part of an array must at least be initialised memory
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (uint32_t i = threadIdx.x; i < Gm::OFF_BAR / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (warp == EPI_WARPS + NGRP) {
        ptx::tmem_alloc(ptx::smem_u32(smem + Gm::OFF_TMEM), 512);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
    const long long n_items = prm.n_items;
    const int n_ph = prm.n_phases;

    if (warp < EPI_WARPS) {
        // ===================================================== epilogue warps
        const int g = NGRP == 2 ? warp >> 2 : 0;
        const int chalf = NGRP == 2 ? 0 : warp >> 2;
        const int wq = warp & 3, wrow = wq * 32;
        const int tid = NGRP == 2 ? (threadIdx.x & 127) : threadIdx.x;
        uint8_t* act = smem + Gm::OFF_ACT + g * Gm::ACT_BYTES;
        const uint32_t tl = tmem_base + ((uint32_t)wrow << 16) + g * (4 * C);
        uint32_t acc_n = 0;
        for (int item = blockIdx.x; item < prm.n_work; item += gridDim.x) {
            const long long i0 = ((long long)item * NGRP + g) * G;
            const int n = (int)max(0LL, min((long long)G, n_items - i0));
            if (n <= 0) continue;
            long long* tr = DBG ? head_trace_slot(prm, item, g, NGRP) : nullptr;
            if (tr && tid == 0) tr[MAX_PHASES * 4 + 0] = clock64();
            load_input<MODE, C>(act, prm, i0, n, tid);
            if (tr && tid == 0) tr[MAX_PHASES * 4 + 1] = clock64();
            ptx::tc_fence_before();
            ptx::fence_proxy_async();
            ptx::mbar_arrive(bar(BAR_ACT + g));
            float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
            for (int ph = 0; ph < n_ph; ++ph) {
                ptx::mbar_wait(bar(BAR_ACC + g), acc_n & 1);
                ++acc_n;
                ptx::tc_fence_after();
                if (tr && tid == 0) tr[ph * 4 + 2] = clock64();
                float* dbg = (DBG && prm.dbg && prm.dbg_phase == ph)
                                 ? prm.dbg + ((long long)item * NGRP + g) * (DBG_ROWS * DBG_COLS) : nullptr;
                if (ph == 0) {
                    epi_conv<MODE, C, ACT, C, 2, Gm::P1, Gm::L, false, false, false, OUT_EO>(
                        act, tl, Gm::B_0, 0, n, Gm::E_ARR, Gm::E_LO, prm, i0, dbg, wrow, lane, chalf, part);
                } else if (ph == 1) {
                    epi_conv<MODE, C, ACT, 2 * C, 1, Gm::P2, Gm::L2, false, false, false, OUT_NAT>(
                        act, tl, Gm::B_1A, 0, n, Gm::S_ARR, Gm::S_LO, prm, i0, dbg, wrow, lane, chalf, part);
                } else if (ph == 2) {
                    epi_conv<MODE, C, ACT, 2 * C, 1, Gm::P2, Gm::L2, true, true, true, OUT_NAT>(
                        act, tl, Gm::B_1B, Gm::B_1S, n, Gm::S_ARR, Gm::S_LO, prm, i0, dbg, wrow, lane, chalf, part);
                } else {
                    const int b = Gm::B_REST + (ph - 3) * 2 * C;
                    if ((ph - 3) % 2 == 0)                  // conv a of a residual block
                        epi_conv<MODE, C, ACT, 2 * C, 1, Gm::P2, Gm::L2, false, false, false, OUT_NAT>(
                            act, tl, b, 0, n, Gm::S_ARR, Gm::S_LO, prm, i0, dbg, wrow, lane, chalf, part);
                    else if (ph + 1 < n_ph)                 // conv b + residual, feeds the next block
                        epi_conv<MODE, C, ACT, 2 * C, 1, Gm::P2, Gm::L2, true, false, true, OUT_NAT>(
                            act, tl, b, 0, n, Gm::S_ARR, Gm::S_LO, prm, i0, dbg, wrow, lane, chalf, part);
                    else                                    // last block: output / pooled linear head
                        epi_conv<MODE, C, ACT, 2 * C, 1, Gm::P2, Gm::L2, true, false, false, OUT_FINAL>(
                            act, tl, b, 0, n, 0, 0, prm, i0, dbg, wrow, lane, chalf, part);
                }
                if (tr && tid == 0) tr[ph * 4 + 3] = clock64();
                if (ph + 1 < n_ph) {
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::mbar_arrive(bar(BAR_ACT + g));
                }
            }
            if (prm.n_out > 0) {
                // AdaptiveAvgPool1d(1) -> Linear (NNTools.py:517-566): per-row dot products were taken above; add the
                // rows of each item in a fixed order.  The exchange buffer aliases the (now idle) operand buffer.
                float* xchg = reinterpret_cast<float*>(act);
                const int m = wrow + lane;
#pragma unroll
                for (int o = 0; o < 4; ++o) xchg[(chalf * 128 + m) * 4 + o] = part[o];
                ptx::named_bar_sync(1 + g, 128 * CS);
                if (tid < n) {
                    float res[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int o = 0; o < prm.n_out; ++o) {
                        float s = 0.f;
                        for (int p = 0; p < Gm::L2; ++p) {
                            float r = xchg[(tid * Gm::P2 + p) * 4 + o];
                            if (CS == 2) r += xchg[(128 + tid * Gm::P2 + p) * 4 + o];
                            s += r;
                        }
                        res[o] = s / (float)Gm::L2 + __ldg(prm.lin_b + o);
                    }
                    if (prm.softmax) {
                        float mx = res[0];
                        for (int o = 1; o < prm.n_out; ++o) mx = fmaxf(mx, res[o]);
                        float sum = 0.f;
                        for (int o = 0; o < prm.n_out; ++o) { res[o] = expf(res[o] - mx); sum += res[o]; }
                        for (int o = 0; o < prm.n_out; ++o) res[o] = res[o] / sum;
                    }
                    for (int o = 0; o < prm.n_out; ++o) prm.out[(i0 + tid) * prm.out_stride + o] = res[o];
                }
                ptx::named_bar_sync(1 + g, 128 * CS);
            }
        }
    } else if (warp < EPI_WARPS + NGRP) {
        // ===================================================== MMA issuers (one warp per group)
        const int g = warp - EPI_WARPS;
        uint32_t slot = 0, par = 0, ar_n = 0;
        const uint32_t act_lo = (ptx::smem_u32(smem + Gm::OFF_ACT) + g * Gm::ACT_BYTES) >> 4;
        const uint32_t ring_lo = ptx::smem_u32(smem + Gm::OFF_W) >> 4;
        const uint32_t d_acc = tmem_base + g * (4 * C);
        for (int item = blockIdx.x; item < prm.n_work; item += gridDim.x) {
            const long long i0 = ((long long)item * NGRP + g) * G;
            const bool active = n_items > i0;
            long long* tr = DBG ? head_trace_slot(prm, item, g, NGRP) : nullptr;
#pragma unroll 1
            for (int ph = 0; ph < n_ph; ++ph) {
                // (no turn-taking between the two compressor groups as in the read convolver: a layer's weights do not
                // fit the ring, and both groups must drain every ring slot before it can be refilled)
                if (active) {
                    ptx::mbar_wait(bar(BAR_ACT + g), ar_n & 1u);
                    ++ar_n;
                    ptx::tc_fence_after();
                }
                if (tr && lane == 0) tr[ph * 4 + 0] = clock64();
                if (ph == 0)
                    issue_phase<MODE, C, 0>(active, act_lo, ring_lo, d_acc, bar(BAR_FULL), bar(BAR_EMPTY), slot, par, lane);
                else if (ph == 1)
                    issue_phase<MODE, C, 1>(active, act_lo, ring_lo, d_acc, bar(BAR_FULL), bar(BAR_EMPTY), slot, par, lane);
                else
                    issue_phase<MODE, C, 2>(active, act_lo, ring_lo, d_acc, bar(BAR_FULL), bar(BAR_EMPTY), slot, par, lane);
                if (active) ptx::tc_commit(bar(BAR_ACC + g));
                if (tr && lane == 0) tr[ph * 4 + 1] = clock64();
                __syncwarp();
            }
        }
    } else {
        // ===================================================== weight producer (one thread, bulk async copies)
        if (lane == 0) {
            uint32_t slot = 0, par = 1;
            const uint32_t w0 = ptx::smem_u32(smem + Gm::OFF_W);
            for (int item = blockIdx.x; item < prm.n_work; item += gridDim.x) {
                {   // pull the next work item's inputs into L2 while this one computes
                    const long long nx = ((long long)(item + gridDim.x) * NGRP) * G;
                    if (nx < n_items) {
                        const long long cnt = min((long long)(NGRP * G), n_items - nx);
                        const uint32_t per = Gm::L * C * 4;
                        ptx::prefetch_l2(prm.in_a + nx * (Gm::L * C), (uint32_t)cnt * per);
                        if (prm.in_s) {
                            const long long s0 = __ldg(prm.site_idx + nx), s1 = __ldg(prm.site_idx + nx + cnt - 1);
                            ptx::prefetch_l2(prm.in_s + s0 * (Gm::L * C), (uint32_t)(s1 - s0 + 1) * per);
                        }
                    }
                }
#pragma unroll 1
                for (int ph = 0; ph < n_ph; ++ph) {
                    const uint32_t total = phase_bytes<MODE, C>(ph);
                    const uint8_t* src = prm.weights + prm.w_src[ph];
                    for (uint32_t o = 0; o < total; o += SLOT_BYTES) {
                        const uint32_t bytes = min(SLOT_BYTES, total - o);
                        ptx::mbar_wait(bar(BAR_EMPTY + slot), par);
                        ptx::mbar_expect_tx(bar(BAR_FULL + slot), bytes);
                        for (uint32_t q = 0; q < bytes; q += 8192u)
                            ptx::bulk_g2s(w0 + slot * SLOT_BYTES + q, src + o + q, min(8192u, bytes - q), bar(BAR_FULL + slot));
                        if (++slot == NSLOT) { slot = 0; par ^= 1u; }
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS + NGRP) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace hc

struct HeadConvTC {
    hc::HeadParams prm;
    uint8_t* d_weights = nullptr;
    float* d_bias = nullptr;       // the pooled linear head (weights, bias)
    int mode = 3, C = 0, sm_count = 148;
    int act = ACT_RELU;            // the one activation of every convolution (Softplus: the [18, 128] expert head only)
    int in_len = 0, out_len = 0, out_ch = 0;
};

// Checks that `net` is "1x1 conv C->C, Res(C->2C, stride 2, 1x1 shortcut), 2 x Res(2C) [, pooled linear]" with
// C = 64 (L = 36) or C = 128 (L = 18) and packs its weights into ring units.
// `covered` receives the number of leading layer records the kernel runs (with up to MAX_EXTRA_BLOCKS appended 2C-channel
// residual blocks, and the pooled linear head when it directly follows them); the caller runs the rest layer by layer.
static HeadConvTC* headconv_tc_create(const std::vector<LayerDesc>& net, int in_len, const float* d_base,
                                      const float* h_base, int precision, std::string& err, size_t* covered = nullptr) {
    using namespace hc;
    if (precision != HELLO_PREC_BF16X3 && precision != HELLO_PREC_BF16) { err = "unknown tensor-core precision"; return nullptr; }
    if (net.size() < 4 || net[0].kind != KIND_CONV) { err = "not a head network"; return nullptr; }
    const int C = net[0].a.cin;
    if (!((C == 64 && in_len == 36) || (C == 128 && in_len == 18))) {
        err = "head network needs [36, 64] or [18, 128] inputs"; return nullptr;
    }
    // ReLU, or Softplus for the expert head of moe_attention_config_single_tech_old_equivalent_layer_norm.py
    const int act = net[0].a.relu;
    if (act != ACT_RELU && !(act == ACT_SOFTPLUS && C == 128)) { err = "head network needs ReLU (or Softplus at 128 channels)"; return nullptr; }
    auto is_res = [&](const LayerDesc& L, int cin, int cout, int s, bool sc) {
        return L.kind == KIND_RES && L.a.cin == cin && L.a.cout == cout && L.a.k == 3 && L.a.stride == s && L.a.pad == 1 &&
               L.a.relu == act && L.b.relu == act && L.b.cin == cout && L.b.cout == cout && L.b.k == 3 && L.b.stride == 1 && L.b.pad == 1 &&
               (L.has_shortcut != 0) == sc && (!sc || (L.s.cin == cin && L.s.cout == cout && L.s.k == 1 && L.s.stride == s && L.s.pad == 0));
    };
    bool ok = net[0].a.cout == C && net[0].a.k == 1 && net[0].a.stride == 1 && net[0].a.pad == 0 &&
              is_res(net[1], C, 2 * C, 2, true) && is_res(net[2], 2 * C, 2 * C, 1, false) && is_res(net[3], 2 * C, 2 * C, 1, false);
    int extra = 0;
    while (ok && extra < MAX_EXTRA_BLOCKS && (size_t)(4 + extra) < net.size() && is_res(net[4 + extra], 2 * C, 2 * C, 1, false)) ++extra;
    const size_t n_blocks_end = 4 + extra;
    const bool pooled = ok && net.size() > n_blocks_end && net[n_blocks_end].kind == KIND_GAP_LINEAR &&
                        net[n_blocks_end].a.cin == 2 * C && net[n_blocks_end].a.cout >= 1 && net[n_blocks_end].a.cout <= 4;
    const size_t n_cov = n_blocks_end + (pooled ? 1 : 0);
    if (covered) *covered = n_cov;
    else ok = ok && net.size() == n_cov;
    if (!ok) { err = "layer table is not conv1x1 / Res(s2) / 2 x Res (/ pooled linear)"; return nullptr; }

    const int parts = precision == HELLO_PREC_BF16X3 ? 2 : 1;
    auto hcv = [&](const ConvDesc& c) { return tc::HostConv{h_base + (c.w - d_base), h_base + (c.b - d_base), c.cin, c.cout, c.k}; };
    std::vector<uint8_t> blob;
    HeadConvTC* t = new HeadConvTC();
    t->mode = parts == 2 ? 3 : 1;
    t->C = C;
    t->act = act;
    t->in_len = C == 64 ? 36 : 18;
    t->out_len = t->in_len / 2;
    t->out_ch = 2 * C;
    std::memset(&t->prm, 0, sizeof(t->prm));

    // unit (tap, j): [hi: 2 chunks][n][8] bf16, then the same for lo; element (chunk c, n, e) = W[n][16j+8c+e][tap]
    auto pack_conv = [&](const tc::HostConv& c) {
        for (int tp = 0; tp < c.k; ++tp)
            for (int j = 0; j < c.cin / 16; ++j) {
                std::vector<uint16_t> hi, lo;
                for (int ch = 0; ch < 2; ++ch)
                    for (int n = 0; n < c.cout; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const float w = c.w[(size_t)(tp * c.cin + 16 * j + 8 * ch + e) * c.cout + n];
                            const uint16_t h = tc::bf16_rne(w);
                            hi.push_back(h);
                            lo.push_back(tc::bf16_rne(w - tc::bf16_to_float(h)));
                        }
                const uint8_t* p = reinterpret_cast<const uint8_t*>(hi.data());
                blob.insert(blob.end(), p, p + hi.size() * 2);
                if (parts == 2) {
                    p = reinterpret_cast<const uint8_t*>(lo.data());
                    blob.insert(blob.end(), p, p + lo.size() * 2);
                }
            }
    };
    std::vector<float> lin(4 * 2 * C + 4, 0.f);               // pooled linear head: weights [n_out][2C], then bias
    auto copy_bias = [&](const tc::HostConv& c, int off) { for (int i = 0; i < c.cout; ++i) t->prm.bias_tab[off + i] = c.b[i]; };
    t->prm.n_phases = N_PHASES + 2 * extra;
    t->prm.w_src[0] = (uint32_t)blob.size(); pack_conv(hcv(net[0].a)); copy_bias(hcv(net[0].a), 0);
    t->prm.w_src[1] = (uint32_t)blob.size(); pack_conv(hcv(net[1].a)); pack_conv(hcv(net[1].s));
    copy_bias(hcv(net[1].a), C); copy_bias(hcv(net[1].s), 3 * C);
    t->prm.w_src[2] = (uint32_t)blob.size(); pack_conv(hcv(net[1].b)); copy_bias(hcv(net[1].b), 5 * C);
    for (int r = 0; r < 2 + extra; ++r) {
        t->prm.w_src[3 + 2 * r] = (uint32_t)blob.size(); pack_conv(hcv(net[2 + r].a)); copy_bias(hcv(net[2 + r].a), 7 * C + (2 * r) * 2 * C);
        t->prm.w_src[4 + 2 * r] = (uint32_t)blob.size(); pack_conv(hcv(net[2 + r].b)); copy_bias(hcv(net[2 + r].b), 7 * C + (2 * r + 1) * 2 * C);
    }
    if (pooled) {
        const ConvDesc& ld = net[n_blocks_end].a;             // linear: w [cout][cin]
        const float* w = h_base + (ld.w - d_base);
        const float* b = h_base + (ld.b - d_base);
        for (int i = 0; i < ld.cout * ld.cin; ++i) lin[i] = w[i];
        for (int i = 0; i < ld.cout; ++i) lin[4 * 2 * C + i] = b[i];
        t->prm.n_out = ld.cout;
    }

    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { err = "cudaGetDeviceProperties failed"; delete t; return nullptr; }
    t->sm_count = prop.multiProcessorCount;
    const size_t smem = C == 64 ? Geo<64>::SMEM_BYTES : Geo<128>::SMEM_BYTES;
    if ((size_t)prop.sharedMemPerBlockOptin < smem) { err = "device has too little shared memory per block"; delete t; return nullptr; }
    if (cudaMalloc(&t->d_weights, blob.size()) != cudaSuccess || cudaMalloc(&t->d_bias, lin.size() * 4) != cudaSuccess ||
        cudaMemcpy(t->d_weights, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t->d_bias, lin.data(), lin.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        err = "allocating the packed bf16 head weights failed";
        if (t->d_weights) cudaFree(t->d_weights);
        if (t->d_bias) cudaFree(t->d_bias);
        delete t;
        return nullptr;
    }
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](const void* fn) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    };
    if (C == 64) {
        if (t->mode == 3) { opt_in((const void*)headconv_tc_kernel<3, 64, false>); opt_in((const void*)headconv_tc_kernel<3, 64, true>); }
        else { opt_in((const void*)headconv_tc_kernel<1, 64, false>); opt_in((const void*)headconv_tc_kernel<1, 64, true>); }
    } else {
        if (t->mode == 3) { opt_in((const void*)headconv_tc_kernel<3, 128, false>); opt_in((const void*)headconv_tc_kernel<3, 128, true>); }
        else { opt_in((const void*)headconv_tc_kernel<1, 128, false>); opt_in((const void*)headconv_tc_kernel<1, 128, true>); }
        if (t->mode == 3) opt_in((const void*)headconv_tc_kernel<3, 128, false, ACT_SOFTPLUS>);
        else opt_in((const void*)headconv_tc_kernel<1, 128, false, ACT_SOFTPLUS>);
    }
    if (e != cudaSuccess) {
        err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        cudaFree(t->d_weights); cudaFree(t->d_bias); delete t;
        return nullptr;
    }
    t->prm.weights = t->d_weights;
    t->prm.lin_w = t->d_bias;
    t->prm.lin_b = t->d_bias + 4 * 2 * C;
    return t;
}

// in_a: fp32 [n, L, C]; in_s/site_idx (optional): operand = 2*in_a - in_s[site_idx].  Feature heads write
// out[n, L/2, 2C]; pooled heads write out[i*out_stride + o] (softmax over o when requested).
static cudaError_t headconv_tc_launch(HeadConvTC* t, const float* in_a, const float* in_s, const int32_t* site_idx,
                                      long long n, float* out, long long out_stride, int softmax, cudaStream_t st,
                                      float* dbg = nullptr, int dbg_phase = -1) {
    if (n <= 0) return cudaSuccess;
    hc::HeadParams prm = t->prm;
    prm.in_a = in_a; prm.in_s = in_s; prm.site_idx = site_idx;
    prm.n_items = n; prm.out = out; prm.out_stride = out_stride; prm.softmax = softmax;
    prm.dbg = dbg; prm.dbg_phase = dbg_phase;
    const int per = t->C == 64 ? hc::Geo<64>::NGRP * hc::Geo<64>::G : hc::Geo<128>::NGRP * hc::Geo<128>::G;
    const long long work = (n + per - 1) / per;
    if (work > 0x7fffffffLL) return cudaErrorInvalidValue;
    prm.n_work = (int)work;
    const int grid = (int)std::min<long long>(work, t->sm_count);
    const bool debug = dbg != nullptr;
#define HELLO_HEAD_LAUNCH(M, CC) \
    do { \
        if (debug) hc::headconv_tc_kernel<M, CC, true><<<grid, hc::Geo<CC>::THREADS, hc::Geo<CC>::SMEM_BYTES, st>>>(prm); \
        else hc::headconv_tc_kernel<M, CC, false><<<grid, hc::Geo<CC>::THREADS, hc::Geo<CC>::SMEM_BYTES, st>>>(prm); \
    } while (0)
    if (t->act == ACT_SOFTPLUS) {                      // 128 channels, no layer dump (the test hook exists for ReLU only)
        using Gs = hc::Geo<128>;
        if (debug) return cudaErrorNotSupported;
        if (t->mode == 3) hc::headconv_tc_kernel<3, 128, false, ACT_SOFTPLUS><<<grid, Gs::THREADS, Gs::SMEM_BYTES, st>>>(prm);
        else hc::headconv_tc_kernel<1, 128, false, ACT_SOFTPLUS><<<grid, Gs::THREADS, Gs::SMEM_BYTES, st>>>(prm);
    } else if (t->C == 64) {
        if (t->mode == 3) HELLO_HEAD_LAUNCH(3, 64); else HELLO_HEAD_LAUNCH(1, 64);
    } else {
        if (t->mode == 3) HELLO_HEAD_LAUNCH(3, 128); else HELLO_HEAD_LAUNCH(1, 128);
    }
#undef HELLO_HEAD_LAUNCH
    return cudaGetLastError();
}

static void headconv_tc_destroy(HeadConvTC* t) {
    if (!t) return;
    if (t->d_weights) cudaFree(t->d_weights);
    if (t->d_bias) cudaFree(t->d_bias);
    delete t;
}

}  // namespace hello
