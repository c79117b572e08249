// Generic fp32 Conv1d (+bias, +ReLU, +residual) as an implicit GEMM on CUDA cores.
//
// Replaces torch.nn.Conv1d called through NNTools.WeightNormedConv1d (reference: python/NNTools.py:791-799) for
// every layer that is not on the fused tcgen05 path: rows = (item, output position), columns = output channels,
// K = tap*cin + ci.  Activations are channel-last [n][L][C] so a GEMM row is contiguous in memory and the output
// tile is a plain row-major store.  Zero padding and item boundaries are handled by predicated loads.
//
// Tile: 128 rows x BN columns x 16 K per step, 256 threads, register tile TM x TN, next K-slab prefetched into
// registers while the current one is consumed from shared memory.
#pragma once
#include "common.cuh"

namespace hello {

struct ConvArgs {
    const void* x;
    long long sn, sl, sc;   // element strides of the input view
    int lin, cin;
    const float* w;         // [K][cout]
    const float* bias;      // [cout]
    float* y;               // [M][cout]
    const float* resid;     // [M][cout] or nullptr; added AFTER the ReLU (NNTools.ResidualBlock.forward, :582-583)
    long long M;            // n_items * lout
    int lout, cout, ksz, stride, pad, relu, K;
};

template <int BN, int TM, int TN, typename TIn, bool VEC>
__global__ void __launch_bounds__(256) conv1d_fp32_kernel(const ConvArgs a) {
    constexpr int BM = 128, BK = 16;
    constexpr int NTX = BN / TN;
    constexpr int NTY = 256 / NTX;
    static_assert(NTY * TM == BM, "thread tile must cover the CTA tile");
    constexpr int EB = BK * BN / 256;   // B-tile floats fetched per thread per K step
    static_assert(EB == 1 || EB == 2 || EB == 4 || EB == 8, "unsupported BN");

    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- A loader: this thread owns row `lrow` of the tile and K offsets [khalf*8, khalf*8+8) of each slab
    const int lrow = tid & (BM - 1);
    const int khalf = tid >> 7;
    const long long lm = m0 + lrow;
    const bool row_ok = lm < a.M;
    long long item = 0;
    int pos0 = 0;
    if (row_ok) {
        item = lm / a.lout;
        pos0 = (int)(lm - item * a.lout) * a.stride - a.pad;
    }
    const TIn* xrow = reinterpret_cast<const TIn*>(a.x) + item * a.sn;

    // ---- B loader: EB consecutive floats of one K row
    const int b_kk = (tid * EB) / BN;
    const int b_nn = (tid * EB) % BN;

    float ra[8];
    float rb[EB];

    auto load_slab = [&](int k0) {
        if (VEC) {
            // cin % 16 == 0 and k0 % 16 == 0: the whole slab lies inside one tap, channels are contiguous
            const int tap = k0 / a.cin;
            const int ci = k0 - tap * a.cin + khalf * 8;
            const int pos = pos0 + tap;
            if (row_ok && pos >= 0 && pos < a.lin) {
                const float4* p = reinterpret_cast<const float4*>(
                    reinterpret_cast<const float*>(xrow) + (long long)pos * a.sl + ci);
                float4 v0 = __ldg(p), v1 = __ldg(p + 1);
                ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
                ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) ra[i] = 0.f;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = k0 + khalf * 8 + i;
                float v = 0.f;
                if (row_ok && k < a.K) {
                    const int tap = k / a.cin;
                    const int ci = k - tap * a.cin;
                    const int pos = pos0 + tap;
                    if (pos >= 0 && pos < a.lin) v = (float)xrow[(long long)pos * a.sl + (long long)ci * a.sc];
                }
                ra[i] = v;
            }
        }
        const int k = k0 + b_kk;
        if (k < a.K) {
            const float* wp = a.w + (long long)k * a.cout + n0 + b_nn;
            if (EB == 8) {
                float4 v0 = __ldg(reinterpret_cast<const float4*>(wp));
                float4 v1 = __ldg(reinterpret_cast<const float4*>(wp) + 1);
                rb[0] = v0.x; rb[1] = v0.y; rb[2] = v0.z; rb[3] = v0.w;
                rb[4 % EB] = v1.x; rb[5 % EB] = v1.y; rb[6 % EB] = v1.z; rb[7 % EB] = v1.w;
            } else if (EB == 4) {
                float4 v0 = __ldg(reinterpret_cast<const float4*>(wp));
                rb[0] = v0.x; rb[1 % EB] = v0.y; rb[2 % EB] = v0.z; rb[3 % EB] = v0.w;
            } else if (EB == 2) {
                float2 v0 = __ldg(reinterpret_cast<const float2*>(wp));
                rb[0] = v0.x; rb[1 % EB] = v0.y;
            } else {
                rb[0] = __ldg(wp);
            }
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) rb[i] = 0.f;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // rows / columns of this thread's register tile (split in halves of 4 so shared loads are conflict-free)
    auto row_of = [&](int i) { return TM == 8 ? (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * TM + i; };
    auto col_of = [&](int j) { return TN == 8 ? (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4)) : tx * TN + j; };

    load_slab(0);
    for (int k0 = 0; k0 < a.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[khalf * 8 + i][lrow] = ra[i];
#pragma unroll
        for (int i = 0; i < EB; ++i) Bs[b_kk][b_nn + i] = rb[i];
        __syncthreads();
        if (k0 + BK < a.K) load_slab(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
            if (TM == 8) {
                float4 v0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                float4 v1 = *reinterpret_cast<const float4*>(&As[kk][BM / 2 + ty * 4]);
                av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
                av[4 % TM] = v1.x; av[5 % TM] = v1.y; av[6 % TM] = v1.z; av[7 % TM] = v1.w;
            } else {
                float4 v0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
            }
            if (TN == 8) {
                float4 v0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                float4 v1 = *reinterpret_cast<const float4*>(&Bs[kk][BN / 2 + tx * 4]);
                bv[0] = v0.x; bv[1] = v0.y; bv[2] = v0.z; bv[3] = v0.w;
                bv[4 % TN] = v1.x; bv[5 % TN] = v1.y; bv[6 % TN] = v1.z; bv[7 % TN] = v1.w;
            } else if (TN == 4) {
                float4 v0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                bv[0] = v0.x; bv[1 % TN] = v0.y; bv[2 % TN] = v0.z; bv[3 % TN] = v0.w;
            } else {
                float2 v0 = *reinterpret_cast<const float2*>(&Bs[kk][tx * 2]);
                bv[0] = v0.x; bv[1 % TN] = v0.y;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias, ReLU, residual add, row-major store
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long m = m0 + row_of(i);
        if (m >= a.M) continue;
        float* yrow = a.y + m * a.cout + n0;
        const float* rrow = a.resid ? a.resid + m * a.cout + n0 : nullptr;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int c = col_of(j);
            float v = acc[i][j] + __ldg(a.bias + n0 + c);
            v = apply_activation(v, a.relu);
            if (rrow) v += __ldg(rrow + c);
            acc[i][j] = v;
        }
        if (TN == 2) {
            *reinterpret_cast<float2*>(yrow + col_of(0)) = make_float2(acc[i][0], acc[i][1 % TN]);
        } else {
            *reinterpret_cast<float4*>(yrow + col_of(0)) =
                make_float4(acc[i][0], acc[i][1 % TN], acc[i][2 % TN], acc[i][3 % TN]);
            if (TN == 8)
                *reinterpret_cast<float4*>(yrow + col_of(4 % TN)) =
                    make_float4(acc[i][4 % TN], acc[i][5 % TN], acc[i][6 % TN], acc[i][7 % TN]);
        }
    }
}

// First convolution of a layer-by-layer model: uint8 pileup rows [item][L][CIN] in, k = 3, stride 1, no padding, so the
// K = 3 * CIN bytes an output needs (18 for six-channel rows) are contiguous in memory.  The implicit-GEMM kernel above pads
// K to two slabs of 16 and spends its time on shared-memory round trips for 18 products; here a thread owns four output
// channels of STEM_ROWS rows, reads its bytes at constant offsets straight from L1 (the lanes of a row read the same
// bytes) and each weight row from shared memory once for all its rows, and a warp store covers whole consecutive output
// rows.  Products are accumulated with fmaf in the same order as conv1d_fp32_kernel (k = tap * cin + ci ascending from
// zero), so the two kernels are bit-identical.
constexpr int STEM_ROWS = 4;

template <int COUT, int CIN>
__global__ void __launch_bounds__(256) conv_stem_u8_kernel(const ConvArgs a) {
    constexpr int K = 3 * CIN;
    constexpr int LPR = COUT / 4;                    // lanes per output row
    constexpr int RPW = 32 / LPR;                    // rows per warp store
    __shared__ __align__(16) float w_s[K * COUT];
    for (int i = threadIdx.x; i < K * COUT; i += 256) w_s[i] = __ldg(a.w + i);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c4 = (lane % LPR) * 4, rw = lane / LPR;
    const long long m0 = (long long)blockIdx.x * (8 * STEM_ROWS * RPW) + (long long)warp * (STEM_ROWS * RPW) + rw;
    float4 acc[STEM_ROWS];
    const uint8_t* src[STEM_ROWS];
#pragma unroll
    for (int r = 0; r < STEM_ROWS; ++r) {
        const long long m = min(m0 + r * RPW, a.M - 1);                        // rows past the end recompute the last one
        const long long item = m / a.lout;
        src[r] = reinterpret_cast<const uint8_t*>(a.x) + item * a.sn + (m - item * a.lout) * CIN;
        acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float4 w = *reinterpret_cast<const float4*>(&w_s[k * COUT + c4]);
#pragma unroll
        for (int r = 0; r < STEM_ROWS; ++r) {
            const float v = (float)__ldg(src[r] + k);
            acc[r].x = fmaf(v, w.x, acc[r].x); acc[r].y = fmaf(v, w.y, acc[r].y);
            acc[r].z = fmaf(v, w.z, acc[r].z); acc[r].w = fmaf(v, w.w, acc[r].w);
        }
    }
    const float4 bias = __ldg(reinterpret_cast<const float4*>(a.bias + c4));
#pragma unroll
    for (int r = 0; r < STEM_ROWS; ++r) {
        const long long m = m0 + r * RPW;
        if (m >= a.M) continue;
        float4 o = make_float4(apply_activation(acc[r].x + bias.x, a.relu), apply_activation(acc[r].y + bias.y, a.relu),
                               apply_activation(acc[r].z + bias.z, a.relu), apply_activation(acc[r].w + bias.w, a.relu));
        *reinterpret_cast<float4*>(a.y + m * COUT + c4) = o;
    }
}

template <int BN, int TM, int TN>
static cudaError_t launch_conv_bn(const ConvArgs& a, bool is_u8, bool vec, cudaStream_t st) {
    dim3 grid((unsigned)((a.M + 127) / 128), (unsigned)(a.cout / BN));
    if (is_u8)
        conv1d_fp32_kernel<BN, TM, TN, uint8_t, false><<<grid, 256, 0, st>>>(a);
    else if (vec)
        conv1d_fp32_kernel<BN, TM, TN, float, true><<<grid, 256, 0, st>>>(a);
    else
        conv1d_fp32_kernel<BN, TM, TN, float, false><<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// y[n][lout][cout] = act(conv(x) + b) (+ resid).  Returns cudaErrorInvalidValue for shapes it does not cover.
static cudaError_t launch_conv(const ActView& x, const ConvDesc& c, long long n_items, float* y,
                               const float* resid, cudaStream_t st) {
    if (n_items <= 0) return cudaSuccess;
    if (c.cout % 16 != 0 || x.ch != c.cin) return cudaErrorInvalidValue;
    ConvArgs a;
    a.x = x.base; a.sn = x.sn; a.sl = x.sl; a.sc = x.sc; a.lin = x.len; a.cin = c.cin;
    a.w = c.w; a.bias = c.b; a.y = y; a.resid = resid;
    a.lout = c.out_len(x.len);
    a.M = n_items * a.lout;
    a.cout = c.cout; a.ksz = c.k; a.stride = c.stride; a.pad = c.pad; a.relu = c.relu; a.K = c.k * c.cin;
    if (x.is_u8 && !resid && c.k == 3 && c.stride == 1 && c.pad == 0 && x.sc == 1 && x.sl == c.cin && (c.cin == 6 || c.cin == 7) &&
        (c.cout == 16 || c.cout == 32) && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(c.b)) & 15) == 0) {
        const int rows_per_block = 8 * STEM_ROWS * (c.cout == 16 ? 8 : 4);
        const unsigned grid = (unsigned)((a.M + rows_per_block - 1) / rows_per_block);
        if (c.cout == 16 && c.cin == 6) conv_stem_u8_kernel<16, 6><<<grid, 256, 0, st>>>(a);
        else if (c.cout == 16) conv_stem_u8_kernel<16, 7><<<grid, 256, 0, st>>>(a);
        else if (c.cin == 6) conv_stem_u8_kernel<32, 6><<<grid, 256, 0, st>>>(a);
        else conv_stem_u8_kernel<32, 7><<<grid, 256, 0, st>>>(a);
        return cudaGetLastError();
    }
    const bool vec = !x.is_u8 && x.sc == 1 && (c.cin % 16 == 0) && (x.sl % 4 == 0) && (x.sn % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(x.base) & 15) == 0);
    if (c.cout % 128 == 0) return launch_conv_bn<128, 8, 8>(a, x.is_u8, vec, st);
    if (c.cout % 64 == 0) return launch_conv_bn<64, 8, 4>(a, x.is_u8, vec, st);
    if (c.cout % 32 == 0) return launch_conv_bn<32, 4, 4>(a, x.is_u8, vec, st);
    return launch_conv_bn<16, 4, 2>(a, x.is_u8, vec, st);
}

}  // namespace hello
