// HBM-bound helper kernels of the MoE forward: pooling, segmented sums over the CSR read->allele and
// allele->site indices, the 2a-s expert input, channel concat, the pooled linear head and the genotype
// posterior epilogue.  All activations are channel-last fp32 [n][L][C].
#pragma once
#include "common.cuh"

namespace hello {

// ---------------------------------------------------------------------------------------------------------
// MaxPool1d(k, stride, pad=0) -- architectures/read_convolver.py:49-56 (144 -> 71).
__global__ void maxpool_kernel(const float4* __restrict__ x, float4* __restrict__ y, long long n_items, int lin,
                               int lout, int c4, int k, int stride) {
    const long long total = n_items * lout * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        const long long r = i / c4;
        const int p = (int)(r % lout);
        const long long n = r / lout;
        const float4* src = x + (n * lin + (long long)p * stride) * c4 + c;
        float4 m = __ldg(src);
        for (int j = 1; j < k; ++j) {
            const float4 v = __ldg(src + (long long)j * c4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
        y[i] = m;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Segmented sum over contiguous row groups (reduceSlots, python/MixtureOfExpertsAdvanced.py:23-34):
// out[g][e] = sum_{r in [off[g], off[g+1])} x[r - row_base][e].  One CTA column-slice per group, float4 lanes,
// rows added in order so the result does not depend on the launch shape.
__global__ void segsum_kernel(const float4* __restrict__ x, float4* __restrict__ out,
                              const int32_t* __restrict__ off, int row_base, int e4) {
    const int g = blockIdx.x;
    const int r0 = off[g] - row_base, r1 = off[g + 1] - row_base;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < e4; e += gridDim.y * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int r = r0;
        for (; r + 4 <= r1; r += 4) {
            const float4 v0 = __ldg(x + (long long)r * e4 + e);
            const float4 v1 = __ldg(x + (long long)(r + 1) * e4 + e);
            const float4 v2 = __ldg(x + (long long)(r + 2) * e4 + e);
            const float4 v3 = __ldg(x + (long long)(r + 3) * e4 + e);
            acc.x = ((acc.x + v0.x) + v1.x) + v2.x + v3.x;
            acc.y = ((acc.y + v0.y) + v1.y) + v2.y + v3.y;
            acc.z = ((acc.z + v0.z) + v1.z) + v2.z + v3.z;
            acc.w = ((acc.w + v0.w) + v1.w) + v2.w + v3.w;
        }
        for (; r < r1; ++r) {
            const float4 v = __ldg(x + (long long)r * e4 + e);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[(long long)g * e4 + e] = acc;
    }
}

// site index of every allele of the chunk: idx[a - a_base] = s (chunk-local)
__global__ void site_index_kernel(const int32_t* __restrict__ site_off, int n_sites, int a_base,
                                  int32_t* __restrict__ idx) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sites) return;
    for (int a = site_off[s] - a_base; a < site_off[s + 1] - a_base; ++a) idx[a] = s;
}

// expert input 2*allele - site (Fork/LinearCombination[2,-1] front of architectures/xattn_subtract.py:13-42):
// the reference evaluates (0 + 2*a) + (-1)*s, i.e. one rounding of 2a - s.
__global__ void two_a_minus_s_kernel(const float4* __restrict__ allele, const float4* __restrict__ site,
                                     const int32_t* __restrict__ site_idx, float4* __restrict__ out,
                                     long long n_alleles, int e4) {
    const long long total = n_alleles * e4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long a = i / e4;
        const int e = (int)(i - a * e4);
        const float4 av = __ldg(allele + i);
        const float4 sv = __ldg(site + (long long)site_idx[a] * e4 + e);
        out[i] = make_float4(fmaf(2.f, av.x, -sv.x), fmaf(2.f, av.y, -sv.y), fmaf(2.f, av.z, -sv.z),
                             fmaf(2.f, av.w, -sv.w));
    }
}

// elementwise sum of two tensors: the legacy wiring's hybrid allele feature `alleleLevelConv0 + alleleLevelConv1`
// (MoEMergedAdvanced.forward with useAdditive and no ConvCombiner, python/MixtureOfExpertsAdvanced.py:408-412)
__global__ void add2_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = __ldg(a + i), y = __ldg(b + i);
        out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
}

// channel concat of two channel-last tensors (ConcatenateChannels, python/NNTools.py:727-733)
__global__ void concat2_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                               long long rows, int ca4, int cb4) {
    const int c4 = ca4 + cb4;
    const long long total = rows * c4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c4;
        const int c = (int)(i - r * c4);
        out[i] = c < ca4 ? __ldg(a + r * ca4 + c) : __ldg(b + r * cb4 + (c - ca4));
    }
}

// ---------------------------------------------------------------------------------------------------------
// terminus: AdaptiveAvgPool1d(1) -> Flatten -> WeightNormedLinear (python/NNTools.py:517-566), optional softmax
// over the outputs (MixtureOfExpertsAdvanced.py:229-232).  One warp per item; warp-shuffle dot products.
__global__ void gap_linear_kernel(const float* __restrict__ x, long long n_items, int len, int cin,
                                  const float* __restrict__ w, const float* __restrict__ b, int cout,
                                  float* __restrict__ out, long long out_stride, int softmax) {
    const long long item = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= n_items) return;
    const float* xi = x + item * (long long)len * cin;
    float res[4];
    for (int o = 0; o < cout; ++o) res[o] = 0.f;
    for (int c = lane; c < cin; c += 32) {
        float s = 0.f;
        for (int p = 0; p < len; ++p) s += __ldg(xi + (long long)p * cin + c);
        const float mean = s / (float)len;
        for (int o = 0; o < cout; ++o) res[o] = fmaf(mean, __ldg(w + (long long)o * cin + c), res[o]);
    }
    for (int o = 0; o < cout; ++o) {
        float v = res[o];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        res[o] = v + __ldg(b + o);
    }
    if (lane == 0) {
        if (softmax) {
            float mx = res[0];
            for (int o = 1; o < cout; ++o) mx = fmaxf(mx, res[o]);
            float sum = 0.f;
            for (int o = 0; o < cout; ++o) { res[o] = expf(res[o] - mx); sum += res[o]; }
            for (int o = 0; o < cout; ++o) res[o] = res[o] / sum;
        }
        for (int o = 0; o < cout; ++o) out[item * out_stride + o] = res[o];
    }
}

__global__ void fill_meta_default_kernel(float* __restrict__ meta, long long n_sites) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_sites) { meta[3 * i] = 1.f; meta[3 * i + 1] = 0.f; meta[3 * i + 2] = 0.f; }
}

// ---------------------------------------------------------------------------------------------------------
// Genotype posterior epilogue: one warp per site.
//   p_e[k]  = sigmoid(logit_e[k])  (absent expert: p = 0)                     MixtureOfExpertsAdvanced.py:530-538
//   P_e[ij] = exp(sum_k log(p t + (1-p)(1-t) + 1e-10)), t = 1 for k in {i,j}   :543-548, 559-571
//   mixed   = (m0*P0 + m1*P1) + m2*P2  in fp32, no FMA contraction             :579-582
//   mix64   = ((0 + P0*m0) + P1*m1) + P2*m2 in float64                         prepareVcf.py:154-158
//   call    = max mixed, ties -> greatest (rank_i, rank_j)                      caller_calling.py:702-705
// Pair order is i <= j, row-major (itertools.product with symmetric dedup).
struct PosteriorArgs {
    const float* logits;        // [3][A_total]; this chunk's alleles start at a_base
    long long logit_stride;     // A_total
    int expert_mask;            // bit e set = expert e present
    const float* meta;          // [S][3] (already filled with (1,0,0) when no meta network)
    const int32_t* site_off;    // global CSR, indexed with global site id
    const int32_t* allele_rank; // [A_total] or nullptr
    const long long* pair_off;  // [S+1]
    long long pair_total;       // P (row stride of pair_prob)
    float* pair_prob;           // [4][P]
    double* pair_mix64;         // [P] or nullptr
    int32_t* best_pair;         // [S][2]
    float* best_prob;           // [S]
    int32_t* call_pair;         // [S][5][2] or nullptr: argmax pair of mixed, e0, e1, e2, float64 re-mix
    double* call_qual;          // [S][5] or nullptr
    int32_t* best_expert;       // [S] or nullptr: np.argmax(meta)
    long long s_begin, s_end;
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// One candidate of the reference's `sorted([(v, k) ...], reverse=True)[0]` (caller_calling.py:702-705,
// prepareVcf.py:59): largest value, ties to the greatest (allele_i, allele_j) key.
struct TopCall {
    double v;
    int i, j, ri, rj;
    __device__ __forceinline__ void offer(double ov, int oi, int oj, int ori, int orj) {
        if (ov > v || (ov == v && (ori > ri || (ori == ri && orj > rj)))) { v = ov; i = oi; j = oj; ri = ori; rj = orj; }
    }
};

// Five calls per site, as the final-call step makes them (prepareVcf.py:36-105, 142-175): 0 the wrapper's fp32
// mixture (what caller_calling.py:702-735 calls), 1-3 each expert on its own, 4 the float64 re-mix ("mean").
// QUAL = -10 log10(1 - min(p, 1 - 1e-8)) in float64 (prepareVcf.py:60-62).
__global__ void posterior_kernel(const PosteriorArgs a) {
    const long long s = a.s_begin + (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= a.s_end) return;
    const int a0 = a.site_off[s];
    const int n = a.site_off[s + 1] - a0;
    const long long p0 = a.pair_off[s];
    const int n_pairs = n * (n + 1) / 2;
    const float m0 = a.meta[3 * s], m1 = a.meta[3 * s + 1], m2 = a.meta[3 * s + 2];

    TopCall top[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) { top[k].v = -1.0; top[k].i = 0; top[k].j = 0; top[k].ri = -1; top[k].rj = -1; }
    for (int q = lane; q < n_pairs; q += 32) {
        // invert q -> (i, j): row i starts at i*n - i*(i-1)/2
        int i = 0, rem = q;
        while (rem >= n - i) { rem -= n - i; ++i; }
        const int j = i + rem;
        float pe[3];
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const bool present = (a.expert_mask >> e) & 1;
            float acc = 0.f;
            for (int k = 0; k < n; ++k) {
                const float p = present ? sigmoid_f(__ldg(a.logits + e * a.logit_stride + a0 + k)) : 0.f;
                const float t = (k == i || k == j) ? 1.f : 0.f;
                const float term = __fadd_rn(__fadd_rn(__fmul_rn(p, t), __fmul_rn(__fsub_rn(1.f, p), __fsub_rn(1.f, t))),
                                             1e-10f);
                acc = __fadd_rn(acc, logf(term));
            }
            pe[e] = expf(acc);
        }
        const float mixed = __fadd_rn(__fadd_rn(__fmul_rn(m0, pe[0]), __fmul_rn(m1, pe[1])), __fmul_rn(m2, pe[2]));
        a.pair_prob[p0 + q] = mixed;
        a.pair_prob[a.pair_total + p0 + q] = pe[0];
        a.pair_prob[2 * a.pair_total + p0 + q] = pe[1];
        a.pair_prob[3 * a.pair_total + p0 + q] = pe[2];
        double d = __dadd_rn(0.0, __dmul_rn((double)pe[0], (double)m0));
        d = __dadd_rn(d, __dmul_rn((double)pe[1], (double)m1));
        d = __dadd_rn(d, __dmul_rn((double)pe[2], (double)m2));
        if (a.pair_mix64) a.pair_mix64[p0 + q] = d;
        const int ri = a.allele_rank ? a.allele_rank[a0 + i] : i;
        const int rj = a.allele_rank ? a.allele_rank[a0 + j] : j;
        top[0].offer((double)mixed, i, j, ri, rj);
#pragma unroll
        for (int e = 0; e < 3; ++e) top[1 + e].offer((double)pe[e], i, j, ri, rj);
        top[4].offer(d, i, j, ri, rj);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, top[k].v, d);
            const int oi = __shfl_xor_sync(0xffffffffu, top[k].i, d);
            const int oj = __shfl_xor_sync(0xffffffffu, top[k].j, d);
            const int ori = __shfl_xor_sync(0xffffffffu, top[k].ri, d);
            const int orj = __shfl_xor_sync(0xffffffffu, top[k].rj, d);
            top[k].offer(ov, oi, oj, ori, orj);
        }
    }
    if (lane == 0) {
        a.best_pair[2 * s] = top[0].i;
        a.best_pair[2 * s + 1] = top[0].j;
        a.best_prob[s] = (float)top[0].v;
        if (a.call_pair) {
#pragma unroll
            for (int k = 0; k < 5; ++k) { a.call_pair[(s * 5 + k) * 2] = top[k].i; a.call_pair[(s * 5 + k) * 2 + 1] = top[k].j; }
        }
        if (a.call_qual) {
#pragma unroll
            for (int k = 0; k < 5; ++k) a.call_qual[s * 5 + k] = -10.0 * log10(1.0 - fmin(top[k].v, 1.0 - 1e-8));
        }
        if (a.best_expert) a.best_expert[s] = (m0 >= m1 && m0 >= m2) ? 0 : (m1 >= m2 ? 1 : 2);   // first maximum
    }
}

}  // namespace hello
