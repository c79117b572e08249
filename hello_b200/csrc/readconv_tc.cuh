// Fused tcgen05 read convolver: the whole 19-convolution per-read stack of HELLO's read_convolver
// (reference: python/architectures/read_convolver.py:9-144, run per read by MoEAttention.forward,
// python/MixtureOfExpertsAdvanced.py:162) in ONE persistent kernel.  uint8 pileup rows go in, fp32 [R,36,64]
// feature maps come out; nothing in between touches HBM.
//
// How a Conv1d becomes tensor-core work
//   Activations live in shared memory channel-chunked: for every group of 8 channels one array of 16-byte rows
//   (row = position, 8 bf16).  That is exactly the K-major "no swizzle" canonical operand layout of tcgen05.mma
//   (core matrix = 8 rows x 16 B, contiguous), with SBO = 128 B and LBO = the distance between chunk arrays.
//   Tap t of a k=3 convolution is then the SAME array read one row further down: the A descriptor of the MMA
//   for tap t simply starts 16 bytes later.  One convolution = taps x (Cin/16) MMAs of shape 128 x Cout x 16
//   accumulating into one TMEM tile; no im2col is ever materialised.
//   Reads are packed back to back along M with a fixed pitch (160 / 80 / 40 rows for the three resolutions);
//   the unused rows of each pitch are kept at zero and double as the zero padding of the pad=1 convolutions.
//   Stride-2 layers (MaxPool1d(3,2), the stride-2 residual block) read an even/odd de-interleaved copy that the
//   previous epilogue writes, which turns stride 2 back into unit row shifts.
//   The 32-channel stage (length 71) is run SPACE-TO-DEPTH: row r of a read holds positions 2r and 2r+1 side by side
//   (64 "channels": [even 32 | odd 32], 36 rows per read, pitch 40 like the 64-channel stage), so a layer is ONE 128-row
//   tile with N = 64 instead of two tiles with N = 32.  The k=3 convolution becomes a centre tap with K = 64 (all four
//   parity blocks non-zero) plus two half taps with K = 32, N = 32 (x[2r-1] only feeds the even output, x[2r+2] only the
//   odd one): no wasted MACs, and the 4 KB activation tile -- the dominant operand cost at small N -- is fetched 16 times
//   per layer instead of 24 (800 instead of 1056 tensor-pipe cycles per group and layer).  The max-pool that opens the
//   stage produces this layout directly: stem conv 3 is evaluated at positions 4r .. 4r+3 into four accumulators per row
//   from a mod-4 de-interleaved copy of the stem conv 2 output.
//
// Pipeline (one CTA per SM, 16 warps, three independent groups of 3 reads in flight)
//   warps 0-11       epilogue warpgroup of read group warp/4 (thread = packed row = TMEM lane): TMEM -> registers
//                    (bias, ReLU, residual, max-pool), split to bf16 hi (+lo), write the next layer's operand back
//                    to shared memory IN PLACE.  The fp32 residual stream stays on chip: half of its 64 columns in
//                    the thread's registers, half in TMEM.
//   warps 12-14      MMA issuer of group warp-12: fully unrolled tcgen05.mma sequences (operand offsets are
//                    immediates, uniform datapath), one elected lane issues.  The groups take turns on the tensor
//                    pipe (token, handed over before a layer's last tap), so one group's layer executes as a block
//                    while the other groups' epilogues run on the CUDA cores.
//   warp 15          one thread streams the weights of the next layer from L2 into a 2-slot ring with
//                    cp.async.bulk (TMA unit) while the current layer computes.
//   Precision: HELLO_PREC_BF16X3 splits activations and weights into bf16 hi+lo (fp32 accumulate, ~2^-17 relative
//   error per product).  tcgen05.mma with both operands in shared memory costs (4096 + 32*N)/128 cycles for a
//   128 x N x 16 tile (measured, tools/mma_bench.cu): the 4 KB activation operand dominates at N = 16..64.  The
//   three products are therefore issued as TWO instructions: A_hi x [W_hi | W_lo] (the weight halves stacked
//   along N, accumulators D_hh | D_hl) and A_lo x W_hi (into D_hh); the epilogue adds the two halves.  That takes
//   128 accumulator columns per group, which is why only three groups (3 x 160 TMEM columns) are in flight and
//   half of the residual moved to registers.  HELLO_PREC_BF16 issues A_hi x W_hi only.
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/hello_moe.h"
#include "common.cuh"
#include "tc_ptx.cuh"

namespace hello {
namespace tc {

constexpr int G = 3;                           // reads per group
constexpr int NG = 3;                          // groups in flight per CTA (one item = NG*G = 9 reads)
constexpr int EW = 4;                          // epilogue warps per group (one per TMEM lane quadrant)
constexpr int P1 = 160, P2 = 80, P3 = 40;      // row pitch of one read at the three resolutions
constexpr int ROWS1 = G * P1, ROWS2 = G * P2, ROWS3 = G * P3;
constexpr int T1 = 4, T2 = 2, T3 = 1;          // 128-row MMA tiles per group
// TMEM columns of a group: [0,128) accumulators, [128,160) the second half of the fp32 residual stream (the first half
// lives in the epilogue threads' registers: 3 x 192 columns would not fit the 512 of an SM)
constexpr uint32_t RES_COL = 128;
constexpr uint32_t GRP_COLS = 160;
constexpr int LIN = 150, LV1 = 148, LV2 = 146, LV3 = 71, LV4 = 36;   // valid lengths
constexpr int LOUT = 36, COUT = 64;
constexpr int N_PHASES = 17;                   // layer phases of the standard read convolver
constexpr int N_RECORDS = 11;                  // its layer records: 3 convs, max-pool, 7 residual blocks
// Addendum models (architectures/read_convolver_addendum.py) append residual blocks at 64 channels: two more phases each,
// same instructions as phases 11..16, so the phase count is a kernel parameter.
constexpr int MAX_EXTRA_BLOCKS = 2;
constexpr int MAX_PHASES = N_PHASES + 2 * MAX_EXTRA_BLOCKS;
constexpr int N_BIAS = 832 + 2 * MAX_EXTRA_BLOCKS * 64;
constexpr int THREADS = (NG * EW + NG + 1) * 32;

// byte layout of one group's activation buffer (offsets relative to its base)
constexpr uint32_t X_STRIDE = (ROWS1 + 8) * 16;   // layer-1 operand: X0[m] = x[m], X1[m] = x[m+1]  (8 ch, hi only)
constexpr uint32_t A1_CH = (ROWS1 + 2) * 16;      // layer-1 output: 2 chunks, natural rows
constexpr uint32_t Q4_ARR = (ROWS3 + 2) * 16;     // layer-2 output de-interleaved mod 4: (chunk, position & 3) arrays, pitch 40
constexpr uint32_t D2_ARR = (ROWS3 + 2) * 16;     // 32-channel stage, space-to-depth: (chunk, parity) arrays, pitch 40, lead zero row
constexpr uint32_t E3_ARR = (ROWS3 + 2) * 16;     // stage-2 output de-interleaved: (chunk, parity), pitch 40, lead row
constexpr uint32_t S3_CH = (ROWS3 + 2) * 16;      // 64-channel stage: 8 chunks, lead zero row
constexpr uint32_t XCHG_OFF = 16 * S3_CH;         // 4 x 32 floats for the max-pool row exchange (first row of every warp slice)
constexpr uint32_t ACT_BYTES = XCHG_OFF + 8 * 128;
constexpr uint32_t WSLOT_BYTES = 49152;

constexpr uint32_t OFF_ACT = 0;
constexpr uint32_t OFF_W = NG * ACT_BYTES;
constexpr uint32_t OFF_BIAS = OFF_W + 2 * WSLOT_BYTES;
constexpr uint32_t OFF_BAR = OFF_BIAS;                 // (the biases live in the kernel parameters)
constexpr uint32_t N_BARS = 4 + 5 * NG;           // w_full[2], w_empty[2], act_ready / acc_full / token / stage_full / stage_free [NG]
constexpr uint32_t OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr uint32_t OFF_STATE = (OFF_TMEM + 16 + 15) & ~15u;   // work range of this CTA + running state of the fused allele sum
constexpr uint32_t OFF_SUM = OFF_STATE + 64;           // fp32 [36][64] running sum of the current allele
// Raw pileup bytes of each group's NEXT work item, copied by the producer thread (cp.async.bulk) while the current item
// computes: the item's first operand is then built from shared memory.  (Built from global memory it took ~6k cycles per
// item -- byte loads queueing behind the epilogues' shared-memory traffic -- with the tensor pipe idle: 8 % of the kernel.)
constexpr uint32_t STG_BYTES = 3712;                   // >= 15 + G * LIN * 8 rounded up to 16, multiple of 128
constexpr uint32_t OFF_STG = (OFF_SUM + LOUT * COUT * 4 + 127) & ~127u;
constexpr uint32_t SMEM_BYTES = OFF_STG + NG * STG_BYTES;
static_assert(SMEM_BYTES <= 232448 && OFF_SUM % 16 == 0 && STG_BYTES >= 16 + G * LIN * 8, "shared memory budget");
static_assert(2 * A1_CH * 2 <= XCHG_OFF && 16 * Q4_ARR <= XCHG_OFF && 16 * D2_ARR <= XCHG_OFF && D2_ARR == E3_ARR &&
              16 * E3_ARR <= XCHG_OFF && 16 * S3_CH <= XCHG_OFF && 2 * X_STRIDE <= XCHG_OFF, "activation layouts");
static_assert(T1 * 128 >= ROWS1 && T2 * 128 >= ROWS2 && T3 * 128 >= ROWS3, "tiles cover the packed rows");
static_assert(T1 * 32 <= RES_COL && T2 * 64 <= RES_COL && T3 * 128 <= RES_COL && RES_COL + 32 <= GRP_COLS &&
              NG * GRP_COLS <= 512, "TMEM budget");

// bias table (floats)
constexpr int B_L1 = 0, B_L2 = 16, B_L3 = 32, B_S2 = 64, B_RCA = 256, B_RCS = 320, B_RCB = 384, B_S3 = 448;

// Per layer phase: where its packed weights live and how many bytes the producer streams into a slot.
struct TcParams {
    uint32_t w_src[MAX_PHASES];
    uint32_t w_bytes[MAX_PHASES];
    int n_phases;             // 17 + 2 per appended residual block
    const uint8_t* reads;
    const uint8_t* weights;
    float bias_tab[N_BIAS];   // biases travel in the kernel parameters: constant-bank loads, no shared-memory bandwidth
    float* out;
    float* dbg;
    long long n_reads;
    int channels, layout, dbg_phase;
    // Fused reads -> alleles sum (reduceSlots, python/MixtureOfExpertsAdvanced.py:23-34 / :163): when allele_out is set the
    // kernel writes, instead of one [36,64] map per read, the sum over each allele's reads -- rows added in read order
    // starting from zero, i.e. exactly what segsum_kernel computes from the per-read maps.
    float* allele_out;            // [n_alleles, 36, 64] or nullptr
    const int32_t* allele_off;    // [n_alleles + 1] read offsets of the alleles (global numbering, minus row_base = local)
    long long n_alleles;
    int row_base;
};

// Packed weights.  One B unit covers one tap and 16 input channels: [2 chunks of 8 channels][rows][8 bf16] with
// rows = the N "hi" weight rows followed (bf16x3 only) by the N "lo" rows.  Units of a phase are stored tap-major.
template <int MODE> __host__ __device__ constexpr uint32_t unit_bytes(int n) { return (MODE == 3 ? 2u : 1u) * n * 32u; }
constexpr int UNITS_L1 = 2, UNITS_L2 = 3, UNITS_L3 = 3, UNITS_RC = 8, UNITS_S3 = 12;
// Space-to-depth 32-channel layer: 4 centre units (input parity x two 16-channel steps) covering both output parities,
// then 2 + 2 half units (tap -1: odd input -> even output; tap +1: even input -> odd output).
constexpr int UNITS_S2_FULL = 4, UNITS_S2_HALF = 4;
template <int MODE> __host__ __device__ constexpr uint32_t s2d_full_bytes() { return (MODE == 3 ? 128u : 64u) * 32u; }
template <int MODE> __host__ __device__ constexpr uint32_t s2d_half_bytes() { return (MODE == 3 ? 64u : 32u) * 32u; }

enum { OUT_NAT = 0, OUT_EO = 1, OUT_GLOBAL = 2, OUT_Q4 = 3 };

// ------------------------------------------------------------------------------------------------ device side
template <int MODE>
__device__ __forceinline__ void store_chunk8(uint8_t* p, uint32_t lo_delta, const float* v) {
    uint32_t h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = ptx::pack_bf16x2(v[2 * q], v[2 * q + 1]);
    *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
    if (MODE == 3) {
        uint32_t l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float d0 = v[2 * q], d1 = v[2 * q + 1];
            ptx::sub2(d0, d1, __uint_as_float(h[q] << 16), __uint_as_float(h[q] & 0xffff0000u));
            l[q] = ptx::pack_bf16x2(d0, d1);
        }
        *reinterpret_cast<uint4*>(p + lo_delta) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// Which bytes the producer stages for the group whose reads start at r0: the 16-byte aligned superset of its rows.
// `ok` is false when that superset would reach outside this launch's read buffer (first / last rows of an unaligned
// buffer); the group then builds its operand from global memory as before.  Producer and consumers call this with the
// same arguments, so they agree.
struct StageDesc { const uint8_t* src; uint32_t off, bytes; bool ok; };
__device__ __forceinline__ StageDesc stage_desc(const uint8_t* reads, long long n_reads_total, int C, long long r0, int n) {
    const long long rb = (long long)LIN * C;
    const uint8_t* p = reads + r0 * rb;
    StageDesc d;
    d.src = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(15));
    d.off = (uint32_t)(p - d.src);
    d.bytes = (d.off + (uint32_t)(n * rb) + 15u) & ~15u;
    d.ok = n > 0 && d.src >= reads && d.src + d.bytes <= reads + n_reads_total * rb;
    return d;
}

// uint8 pileup rows of one group -> layer-1 operand (two row-shifted copies, bf16, channels padded to 8), in two steps so
// that the bytes of the NEXT item can be fetched into registers before the current item's output is summed (the fetch then
// overlaps the wait for this group's turn in the allele sum), and stored once the activation buffer is free.
constexpr int IN_IT = (ROWS1 + 8 + EW * 32 - 1) / (EW * 32);      // rows per thread
// STAGED: `reads` + r0 rows is the group's staging buffer in shared memory instead of global memory.
template <bool STAGED>
__device__ __forceinline__ void fetch_rows(uint32_t (&raw)[IN_IT][2], const uint8_t* __restrict__ reads, long long r0,
                                           int n_reads, int C, int layout, int tid) {
    // all loads of the thread's rows are issued before the first one is used
#pragma unroll
    for (int it = 0; it < IN_IT; ++it) {
        const int m = tid + it * (EW * 32);
        const int i = m / P1, p = m - i * P1;
        const bool ok = m < ROWS1 + 8 && i < n_reads && p < LIN;
        const uint8_t* src = reads + (r0 + (ok ? i : 0)) * (long long)(LIN * C);
        uint32_t b[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            uint32_t x = 0u;
            if (ok && ch < C) {
                const uint8_t* q = layout == HELLO_LAYOUT_RLC ? src + p * C + ch : src + ch * LIN + p;
                x = STAGED ? *q : __ldg(q);
            }
            b[ch] = x;
        }
        raw[it][0] = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);     // kept packed: 2 registers per row
        raw[it][1] = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    }
}

__device__ __forceinline__ void store_rows(uint8_t* act, const uint32_t (&raw)[IN_IT][2], int tid) {
#pragma unroll
    for (int it = 0; it < IN_IT; ++it) {
        const int m = tid + it * (EW * 32);
        if (m < ROWS1 + 8) {
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {                             // 0..255 are exact in bf16
                const uint32_t two = raw[it][q >> 1] >> (16 * (q & 1));
                w[q] = (__float_as_uint((float)(two & 0xffu)) >> 16) | (__float_as_uint((float)((two >> 8) & 0xffu)) & 0xffff0000u);
            }
            const uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(act + (uint32_t)m * 16) = v;
            if (m >= 1) *reinterpret_cast<uint4*>(act + X_STRIDE + (uint32_t)(m - 1) * 16) = v;
        }
    }
    if (tid == 0) *reinterpret_cast<uint4*>(act + X_STRIDE + (uint32_t)(ROWS1 + 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
}

// Epilogue of one convolution for one group.  Thread = row (TMEM lane).  The accumulator tiles are cut into four
// blocks of 16 columns (TILES * N / 16 == 4 for every layer).  The fp32 residual of blocks 0-1 lives in `rr`
// (registers), that of blocks 2-3 in TMEM.
//   y = relu(acc_hh [+ acc_hl] + bias) [+ resid (+ bias2)], invalid rows forced to zero, written as the next operand.
//   STACK: the layer was issued in the stacked form (two accumulator halves per tile).
//   MOVE_SC (stride-2 block): the shortcut accumulators sit next to conv a's in columns [64,128); move them to
//   the residual storage before the next layer reuses the accumulator columns.
//   S2D (32-channel stage, space-to-depth): the row holds two positions, block = (parity, 16-channel half); the
//   accumulator columns are [hl even | hh even | hh odd | hl odd] (bf16x3) or [even | odd] (bf16), see issue_stage2_s2d;
//   LVALID counts rows of the even positions, the odd ones have one fewer.
//   OUT_Q4 (stem conv 2): output rows de-interleaved mod 4 into (chunk, position & 3) arrays of pitch P3.
template <int MODE, int ACT, bool STACK, int N, int TILES, int PITCH, int LVALID, bool RESID, bool RES_BIAS, bool WRITE_RESID,
          bool MOVE_SC, int OUT, int LEAD, bool S2D = false>
__device__ __forceinline__ void epi_conv(const TcParams& prm, uint8_t* act, uint32_t tl, int bias, int bias2, int n_reads,
                                         uint32_t out_stride, uint32_t out_lo, float* __restrict__ gout,
                                         float* __restrict__ dbg, int wrow, int lane, float (&rr)[32],
                                         long long* tr = nullptr) {
    static_assert(TILES * N == 64, "four 16-column blocks per layer");
    constexpr int ROWS = G * PITCH;
    constexpr int BPT = N / 16;                       // blocks per tile
    constexpr bool TWO = MODE == 3 && STACK;
    constexpr uint32_t TILE_COLS = TWO ? 2 * N : N;
    // One block = 16 accumulator columns of one tile.  REG = the block's residual lives in registers rr[RO .. RO+16)
    // (blocks 0 and 1), otherwise in TMEM.  The body is instantiated as rarely as the register indexing allows and the
    // rest is a rolled loop: the kernel's code size is what the instruction cache sees (the epilogues were stalling
    // on instruction fetch), not its instruction count.
    auto block = [&](int blk, auto reg_tag, auto ro_tag) {
        constexpr bool REG = decltype(reg_tag)::value;
        constexpr int RO = decltype(ro_tag)::value;
        const int tile = S2D ? 0 : blk / BPT, c0 = S2D ? (blk & 1) * 16 : (blk - tile * BPT) * 16;
        const int par = S2D ? (blk >> 1) : 0;
        const int m = tile * 128 + wrow + lane;
        const int i = m / PITCH, p = m - i * PITCH;
        const bool valid = (i < n_reads) && (p < LVALID - par);
        const bool in_buf = m < ROWS;
        float x[16];
        float w[TWO ? 16 : 1];
        float r[((RESID && !REG) || MOVE_SC) ? 16 : 1];
        const uint32_t xcol = S2D ? (TWO ? (par ? 64u : 32u) : par * 32u) + c0 : tile * TILE_COLS + c0;
        const uint32_t wcol = S2D ? (par ? 96u : 0u) + c0 : tile * TILE_COLS + N + c0;
        ptx::tmem_ld16(tl + xcol, x);
        if (TWO) ptx::tmem_ld16(tl + wcol, w);
        if (RESID && !REG) ptx::tmem_ld16(tl + RES_COL + (blk - 2) * 16, r);
        if (MOVE_SC) ptx::tmem_ld16(tl + 64 + c0, r);
        ptx::tmem_wait_ld();
        if (tr && blk == 0 && wrow + lane == 0) tr[4] = clock64();
#pragma unroll
        for (int c = 0; c < 16; c += 2) {               // packed fp32x2 adds: same roundings, half the issue slots
            float y0 = x[c], y1 = x[c + 1];
            if (TWO) ptx::add2(y0, y1, w[c], w[c + 1]);
            ptx::add2(y0, y1, prm.bias_tab[bias + c0 + c], prm.bias_tab[bias + c0 + c + 1]);
            if (ACT == ACT_RELU) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
            else { y0 = softplus_fast(y0); y1 = softplus_fast(y1); }
            if (RESID) {
                float r0 = REG ? rr[RO + c] : r[c], r1 = REG ? rr[RO + c + 1] : r[c + 1];
                if (RES_BIAS) ptx::add2(r0, r1, prm.bias_tab[bias2 + c0 + c], prm.bias_tab[bias2 + c0 + c + 1]);
                ptx::add2(y0, y1, r0, r1);
            }
            x[c] = valid ? y0 : 0.f;
            x[c + 1] = valid ? y1 : 0.f;
        }
        if (WRITE_RESID || MOVE_SC) {
            if (REG) {
#pragma unroll
                for (int c = 0; c < 16; ++c) rr[RO + c] = MOVE_SC ? r[c] : x[c];
            } else {
                ptx::tmem_st16(tl + RES_COL + (blk - 2) * 16, MOVE_SC ? r : x);
            }
        }
        if (dbg) {
#pragma unroll
            for (int c = 0; c < 16; ++c) dbg[m * 64 + par * 32 + c0 + c] = x[c];
        }
        if (OUT == OUT_GLOBAL) {
            // Stage the fp32 row in the (now idle) operand buffer, 16-byte chunks XOR-swizzled by the row so that the 32
            // rows of a warp hit all banks; the group then copies whole reads out with fully coalesced stores
            // (a thread storing its own row straight to HBM touches 32 lines per instruction and took 2.5x longer).
            if (in_buf) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (c0 / 4 + q) ^ (m & 15);
                    *reinterpret_cast<float4*>(act + (uint32_t)m * 256 + chunk * 16) =
                        make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                }
            }
        } else if (in_buf) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c8 = c0 / 8 + q;
                uint8_t* dst = S2D ? act + (c8 * 2 + par) * out_stride + (uint32_t)(m + LEAD) * 16
                               : OUT == OUT_NAT ? act + c8 * out_stride + (uint32_t)(m + LEAD) * 16
                               : OUT == OUT_Q4 ? act + (c8 * 4 + (p & 3)) * out_stride + (uint32_t)(i * P3 + (p >> 2)) * 16
                                               : act + (c8 * 2 + (m & 1)) * out_stride + (uint32_t)((m >> 1) + LEAD) * 16;
                store_chunk8<MODE>(dst, out_lo, x + 8 * q);
            }
        }
        if (tr && wrow + lane == 0 && (blk == 0 || blk == 3)) tr[blk == 0 ? 5 : 6] = clock64();
    };
    using T = std::true_type;
    using F = std::false_type;
    if (RESID || WRITE_RESID || MOVE_SC) {
        block(0, T{}, std::integral_constant<int, 0>{});
        block(1, T{}, std::integral_constant<int, 16>{});
#pragma unroll 1
        for (int blk = 2; blk < 4; ++blk) block(blk, F{}, std::integral_constant<int, 0>{});
    } else {
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) block(blk, F{}, std::integral_constant<int, 0>{});
    }
    if (LEAD && OUT != OUT_GLOBAL && wrow + lane == 0) {          // zero padding row in front of the first read
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        constexpr int ARRS = (N / 8) * (OUT == OUT_EO ? 2 : 1);
#pragma unroll
        for (int a = 0; a < ARRS; ++a) {
            *reinterpret_cast<uint4*>(act + a * out_stride) = z;
            if (MODE == 3) *reinterpret_cast<uint4*>(act + a * out_stride + out_lo) = z;
        }
    }
    if (WRITE_RESID || MOVE_SC) ptx::tmem_wait_st();
}

// Copy the staged [n_reads x 36 x 64] fp32 features of a group to HBM: consecutive threads store consecutive 16 bytes.
__device__ __forceinline__ void copy_out(const uint8_t* act, float* __restrict__ gout, int n_reads, int g, int tid) {
    ptx::named_bar_sync(1 + g, EW * 32);
    const int total = n_reads * LOUT * (COUT / 4);
    float4* dst = reinterpret_cast<float4*>(gout);
    for (int f = tid; f < total; f += EW * 32) {
        const int row = f >> 4, q = f & 15;
        const int i = row / LOUT, p = row - i * LOUT;
        const int m = i * P3 + p;
        dst[f] = *reinterpret_cast<const float4*>(act + (uint32_t)m * 256 + ((q ^ (m & 15)) * 16));
    }
    ptx::named_bar_sync(1 + g, EW * 32);              // the buffer is free for the next work item's input
}

// This CTA's share of the reads, state of the fused allele sum (shared memory, OFF_STATE).
struct CtaState {
    long long r_begin, r_end;     // reads [r_begin, r_end) (chunk-local numbering)
    long long cur_end;            // first read after the allele being summed
    int cur_allele;               // allele being summed (chunk-local index)
    volatile int turn;            // next group (in read order) allowed to add its reads
};

// Fused reduceSlots: the groups of a CTA hold consecutive reads and a CTA owns whole alleles, so adding every group's
// staged rows, in read order, into one running [36,64] sum and flushing it when the allele changes reproduces
// segsum_kernel bit for bit (same additions in the same order) without the per-read maps ever reaching HBM.
__device__ __forceinline__ void accumulate_out(uint8_t* smem, const uint8_t* act, const TcParams& prm, long long r0, int n_reads,
                                               int seq, int g, int tid) {
    CtaState* st = reinterpret_cast<CtaState*>(smem + OFF_STATE);
    float4* sum = reinterpret_cast<float4*>(smem + OFF_SUM);
    if (tid == 0) { while (st->turn != seq) __nanosleep(32); }             // our turn: earlier reads are in the sum
    ptx::named_bar_sync(1 + g, EW * 32);                                   // ... and the group's rows are staged
    int cur = st->cur_allele;
    long long cur_end = st->cur_end;
    constexpr int SLOTS = LOUT * COUT / 4;                                 // float4 slots of one map
    for (int i = 0; i < n_reads; ++i) {
        if (r0 + i == cur_end) {                                           // allele complete: flush, start the next one
            float4* dst = reinterpret_cast<float4*>(prm.allele_out + (long long)cur * (LOUT * COUT));
            for (int f = tid; f < SLOTS; f += EW * 32) { dst[f] = sum[f]; sum[f] = make_float4(0.f, 0.f, 0.f, 0.f); }
            ++cur;
            cur_end = (long long)__ldg(prm.allele_off + cur + 1) - prm.row_base;
        }
        for (int f = tid; f < SLOTS; f += EW * 32) {
            const int p = f >> 4, q = f & 15, m = i * P3 + p;
            const float4 v = *reinterpret_cast<const float4*>(act + (uint32_t)m * 256 + ((q ^ (m & 15)) * 16));
            float4 a = sum[f];
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            sum[f] = a;
        }
    }
    if (r0 + n_reads == st->r_end) {                                       // last reads of this CTA: flush the last allele
        float4* dst = reinterpret_cast<float4*>(prm.allele_out + (long long)cur * (LOUT * COUT));
        for (int f = tid; f < SLOTS; f += EW * 32) dst[f] = sum[f];
    }
    ptx::named_bar_sync(1 + g, EW * 32);                                   // sums written; the staging buffer is free
    if (tid == 0) {
        st->cur_allele = cur;
        st->cur_end = cur_end;
        __threadfence_block();
        st->turn = seq + 1;
    }
}

// Epilogue of stem conv 3 fused with MaxPool1d(3,2), producing the space-to-depth layout of the 32-channel stage.  Row r of
// a read holds the convolution at positions 4r .. 4r+3 in four accumulators a0..a3 (32 columns each, the sum of their three
// operand products); pooled[2r] = max(a0, a1, a2), pooled[2r+1] = max(a2, a3, a0 of row r+1) -- the latter comes from the
// neighbouring lane (shared-memory exchange across warp borders).  Starts the residual stream: the even position's 32
// channels -> registers, the odd position's -> TMEM (blocks 0,1 / 2,3 of epi_conv<S2D>).
template <int MODE, int ACT>
__device__ __forceinline__ void epi_pool(const TcParams& prm, uint8_t* act, uint32_t tl, int bias, int n_reads, int g, int wq,
                                         int lane, float* __restrict__ dbg, float (&rr)[32]) {
    float* xchg = reinterpret_cast<float*>(act + XCHG_OFF);
    {
        float e[32];
        ptx::tmem_ld32(tl, e);
        ptx::tmem_wait_ld();
        if (lane == 0) {
            float4* dst = reinterpret_cast<float4*>(xchg + wq * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(e[4 * q], e[4 * q + 1], e[4 * q + 2], e[4 * q + 3]);
        }
    }
    ptx::named_bar_sync(1 + g, EW * 32);
    const int m = wq * 32 + lane;
    const int i = m / P3, p = m - i * P3;
    const bool valid_e = (i < n_reads) && (p < LV4), valid_o = (i < n_reads) && (p < LV4 - 1);
    // row m+1 of the last lane lives in the next warp slice: lane c fetches its column-c value once and broadcasts it
    // (the very last row of the group is padding, any finite value will do)
    const float xv = xchg[((wq + 1) & 3) * 32 + lane];
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
        float ev[16], ov[16];
        {
            float a0[16], a1[16];
            ptx::tmem_ld16(tl + hb * 16, a0);
            ptx::tmem_ld16(tl + 32 + hb * 16, a1);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float dn = __shfl_down_sync(0xffffffffu, a0[c], 1);
                const float nb = __shfl_sync(0xffffffffu, xv, hb * 16 + c);
                ov[c] = lane == 31 ? nb : dn;                    // a0 of row m+1
                ev[c] = fmaxf(a0[c], a1[c]);
            }
        }
        {
            float a2[16], a3[16];
            ptx::tmem_ld16(tl + 64 + hb * 16, a2);
            ptx::tmem_ld16(tl + 96 + hb * 16, a3);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float bb = prm.bias_tab[bias + hb * 16 + c];
                // the activation is monotone, so it commutes with the max: one evaluation per pooled value
                const float em = fmaxf(ev[c], a2[c]) + bb, om = fmaxf(fmaxf(a2[c], a3[c]), ov[c]) + bb;
                const float e = ACT == ACT_RELU ? fmaxf(em, 0.f) : softplus_fast(em);
                const float o = ACT == ACT_RELU ? fmaxf(om, 0.f) : softplus_fast(om);
                ev[c] = valid_e ? e : 0.f;
                ov[c] = valid_o ? o : 0.f;
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) rr[hb * 16 + c] = ev[c];
        ptx::tmem_st16(tl + RES_COL + hb * 16, ov);
        if (dbg) {
#pragma unroll
            for (int c = 0; c < 16; ++c) { dbg[m * 64 + hb * 16 + c] = ev[c]; dbg[m * 64 + 32 + hb * 16 + c] = ov[c]; }
        }
        if (m < ROWS3) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                store_chunk8<MODE>(act + ((hb * 2 + q) * 2 + 0) * D2_ARR + (uint32_t)(m + 1) * 16, 8 * D2_ARR, ev + 8 * q);
                store_chunk8<MODE>(act + ((hb * 2 + q) * 2 + 1) * D2_ARR + (uint32_t)(m + 1) * 16, 8 * D2_ARR, ov + 8 * q);
            }
        }
    }
    if (wq == 0 && lane == 0) {                                   // zero padding row in front of the first read
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            *reinterpret_cast<uint4*>(act + a * D2_ARR) = z;
            if (MODE == 3) *reinterpret_cast<uint4*>(act + a * D2_ARR + 8 * D2_ARR) = z;
        }
    }
    ptx::tmem_wait_st();
}

// All tcgen05.mma of one 128-row tile of one convolution, fully unrolled: every operand offset is an immediate,
// so one MMA costs two integer adds plus the issue.  Runs on a converged warp with warp-uniform values.
//   AO0..2  byte offset of the A operand of each tap (row shift / even-odd array selection)
//   A_LBO   distance between 8-channel chunk arrays;  A_LO  distance to the "lo" copy of the activations
//   STACK   bf16x3 as two instructions: A_hi x [W_hi | W_lo] -> D[0,2N), A_lo x W_hi -> D[0,N);
//           otherwise three instructions into D[0,N) (used where the accumulator columns are scarce)
//   PART 0: all but the last few (tap, 16-channel) steps;  PART 1: those last steps.  The issuer hands the tensor-pipe
//   token to the next group between the two parts, so the hand-over latency hides behind this group's last MMAs
//   (a longer tail only interleaves the two groups in the pipe's FIFO and delays this group's accumulators).
template <int MODE, int PART, bool STACK, int N, int NTAPS, int K16, bool A_HAS_LO, uint32_t AO0, uint32_t AO1,
          uint32_t AO2, uint32_t A_LBO, uint32_t A_LO, uint32_t B_UNIT0>
__device__ __forceinline__ void issue_tile(uint32_t act_lo, uint32_t w_lo, uint32_t d) {
    constexpr uint32_t AO[3] = {AO0 >> 4, AO1 >> 4, AO2 >> 4};
    constexpr uint32_t BROWS = MODE == 3 ? 2 * N : N;
    constexpr uint32_t UNIT = BROWS * 32;
    constexpr uint32_t idesc_n = ptx::idesc_bf16_m128(N);
    constexpr uint32_t idesc_2n = ptx::idesc_bf16_m128(2 * N);
    // PART 1 = the last TAIL (tap, 16-channel) steps: ~4 MMAs per group, enough to cover the token hand-over
    constexpr int STEPS = NTAPS * K16;
    constexpr int TAIL = NTAPS == 1 ? STEPS : (K16 >= 4 ? 2 : 1);
    constexpr int S_BEGIN = PART == 0 ? 0 : STEPS - TAIL, S_END = PART == 0 ? STEPS - TAIL : STEPS;
    const uint32_t a0 = act_lo | (((A_LBO >> 4) & 0x3FFFu) << 16);
    const uint32_t b0 = (w_lo + ((B_UNIT0 * UNIT) >> 4)) | (((BROWS * 16u) >> 4) << 16);
#pragma unroll
    for (int s = S_BEGIN; s < S_END; ++s) {
        {
            const int t = s / K16, j = s - t * K16;
            const uint32_t a = a0 + AO[t] + (uint32_t)j * ((2u * A_LBO) >> 4);
            const uint32_t b = b0 + (uint32_t)(t * K16 + j) * (UNIT >> 4);
            const uint32_t acc = (t == 0 && j == 0) ? 0u : 1u;
            if (MODE != 3) {
                ptx::mma_bf16_ss(d, a, b, idesc_n, acc);
            } else if (STACK) {
                ptx::mma_bf16_ss(d, a, b, idesc_2n, acc);
                if (A_HAS_LO) ptx::mma_bf16_ss(d, a + (A_LO >> 4), b, idesc_n, 1u);
            } else {
                ptx::mma_bf16_ss(d, a, b, idesc_n, acc);
                ptx::mma_bf16_ss(d, a, b + ((N * 16u) >> 4), idesc_n, 1u);
                if (A_HAS_LO) ptx::mma_bf16_ss(d, a + (A_LO >> 4), b, idesc_n, 1u);
            }
        }
    }
}

// One layer of the 32-channel stage in its space-to-depth form (see the file header).  Activation arrays: (chunk c8,
// parity) at (c8 * 2 + parity) * D2_ARR, lead zero row, "lo" copies 8 arrays further; a 16-channel K step of one parity is
// the array pair (4 j + parity, 4 j + 2 + parity), i.e. LBO = 2 * D2_ARR and 4 * D2_ARR between steps.
// Weight units in issue order (packed by pack_s2d_phase): 2 x centre / even input, 2 x centre / odd input (128 rows
// [lo even | hi even | hi odd | lo odd]; the A_lo x W_hi product reads rows [32,96) of the same unit), 2 x tap -1 (odd input
// of the row above -> even output, 64 rows [lo even | hi even]), 2 x tap +1 (even input of the row below -> odd output, 64
// rows [hi odd | lo odd]).  Accumulator columns: [hl even | hh even | hh odd | hl odd].  bf16 (MODE 1): units hold the hi rows
// only, accumulators [even | odd].  PART 1 = the two tap +1 steps (token hand-over tail).
template <int MODE, int PART>
__device__ __forceinline__ void issue_stage2_s2d(uint32_t act_lo, uint32_t w_lo, uint32_t d) {
    constexpr uint32_t ARR = D2_ARR >> 4, A_LO = (8 * D2_ARR) >> 4;
    constexpr uint32_t FULL = s2d_full_bytes<MODE>() >> 4, HALF = s2d_half_bytes<MODE>() >> 4;
    constexpr uint32_t FROWS = MODE == 3 ? 128 : 64, HROWS = MODE == 3 ? 64 : 32;
    const uint32_t a0 = act_lo | ((((2 * D2_ARR) >> 4) & 0x3FFFu) << 16);
    const uint32_t bf = w_lo | (((FROWS * 16u) >> 4) << 16);            // full units: chunk distance = 128 (64) rows
    const uint32_t bh = (w_lo + 4 * FULL) | (((HROWS * 16u) >> 4) << 16);
    constexpr uint32_t i128 = ptx::idesc_bf16_m128(128), i64 = ptx::idesc_bf16_m128(64), i32 = ptx::idesc_bf16_m128(32);
    if (PART == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {                                   // centre: row m of the tile = array row m + 1
            const uint32_t par = u >> 1, j = u & 1;
            const uint32_t a = a0 + par * ARR + 1u + j * 4 * ARR;
            const uint32_t b = bf + u * FULL;
            if (MODE == 3) {
                ptx::mma_bf16_ss(d, a, b, i128, u == 0 ? 0u : 1u);
                ptx::mma_bf16_ss(d + 32, a + A_LO, b + 32, i64, 1u);     // A_lo x [hi even | hi odd]
            } else {
                ptx::mma_bf16_ss(d, a, b, i64, u == 0 ? 0u : 1u);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {                                   // tap -1: odd input of array row m -> even output
            const uint32_t a = a0 + ARR + j * 4 * ARR;
            const uint32_t b = bh + j * HALF;
            if (MODE == 3) {
                ptx::mma_bf16_ss(d, a, b, i64, 1u);
                ptx::mma_bf16_ss(d + 32, a + A_LO, b + 32, i32, 1u);
            } else {
                ptx::mma_bf16_ss(d, a, b, i32, 1u);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {                                   // tap +1: even input of array row m + 2 -> odd output
            const uint32_t a = a0 + 2u + j * 4 * ARR;
            const uint32_t b = bh + (2 + j) * HALF;
            if (MODE == 3) {
                ptx::mma_bf16_ss(d + 64, a, b, i64, 1u);
                ptx::mma_bf16_ss(d + 64, a + A_LO, b, i32, 1u);
            } else {
                ptx::mma_bf16_ss(d + 32, a, b, i32, 1u);
            }
        }
    }
}

// All MMAs of layer phase `ph` for one group.  act_lo / w_lo: shared-memory addresses >> 4; d0: TMEM base of the group.
template <int MODE, int PART>
__device__ __forceinline__ void issue_phase(int ph, uint32_t act_lo, uint32_t w_lo, uint32_t d0) {
    constexpr uint32_t TC16 = MODE == 3 ? 32 : 16;   // accumulator columns per tile of the 16-channel stem layers
    if (ph == 0) {
#pragma unroll 1
        for (uint32_t t = 0; t < T1; ++t)
            issue_tile<MODE, PART, true, 16, 2, 1, false, 0, 32, 0, X_STRIDE, 0, 0>(act_lo + t * 128u, w_lo, d0 + t * TC16);
    } else if (ph == 1) {
#pragma unroll 1
        for (uint32_t t = 0; t < T1; ++t)
            issue_tile<MODE, PART, true, 16, 3, 1, true, 0, 16, 32, A1_CH, 2 * A1_CH, 0>(act_lo + t * 128u, w_lo, d0 + t * TC16);
    } else if (ph == 2) {
        // Stem conv 3 at positions 4r + j (j = 0..3), one accumulator of 32 columns each, from the mod-4 de-interleaved
        // layer-2 output: x[4r + j + t] sits in array (j + t) & 3 at row r + ((j + t) >> 2).  Four accumulators fill the
        // group's columns, so the three products go into the same 32 columns.
        issue_tile<MODE, PART, false, 32, 3, 1, true, 0, Q4_ARR, 2 * Q4_ARR, 4 * Q4_ARR, 8 * Q4_ARR, 0>(act_lo, w_lo, d0);
        issue_tile<MODE, PART, false, 32, 3, 1, true, Q4_ARR, 2 * Q4_ARR, 3 * Q4_ARR, 4 * Q4_ARR, 8 * Q4_ARR, 0>(act_lo, w_lo, d0 + 32u);
        issue_tile<MODE, PART, false, 32, 3, 1, true, 2 * Q4_ARR, 3 * Q4_ARR, 16, 4 * Q4_ARR, 8 * Q4_ARR, 0>(act_lo, w_lo, d0 + 64u);
        issue_tile<MODE, PART, false, 32, 3, 1, true, 3 * Q4_ARR, 16, Q4_ARR + 16, 4 * Q4_ARR, 8 * Q4_ARR, 0>(act_lo, w_lo, d0 + 96u);
    } else if (ph < 9) {
        issue_stage2_s2d<MODE, PART>(act_lo, w_lo, d0);
    } else if (ph == 9) {
        // stride 2: x[2p-1], x[2p], x[2p+1] = odd[p-1], even[p], odd[p]; the 1x1 shortcut reads even[p].  Both results
        // must sit in the accumulator columns at once (conv a in [0,64), shortcut in [64,128)): unstacked form.
        issue_tile<MODE, PART, false, 64, 3, 2, true, E3_ARR, 16, E3_ARR + 16, 2 * E3_ARR, 8 * E3_ARR, 0>(act_lo, w_lo, d0);
        issue_tile<MODE, PART, false, 64, 1, 2, true, 16, 0, 0, 2 * E3_ARR, 8 * E3_ARR, 6>(act_lo, w_lo, d0 + 64u);
    } else {
        issue_tile<MODE, PART, true, 64, 3, 4, true, 0, 16, 32, S3_CH, 8 * S3_CH, 0>(act_lo, w_lo, d0);
    }
}

// Timeline hook (dbg_phase == TRACE_PHASE): CTA 0 stamps clock64() for its first TRACE_ITEMS work items into the debug
// buffer as int64 [item][group][phase][8] = {issue start, issue end, accumulators seen by the epilogue, epilogue end,
// first TMEM load landed, first block stored, last block stored, after the operand fence}.
constexpr int TRACE_PHASE = -2, TRACE_ITEMS = 16;
__device__ __forceinline__ long long* trace_slot(const TcParams& prm, int item, int g) {
    if (!prm.dbg || prm.dbg_phase != TRACE_PHASE || blockIdx.x != 0) return nullptr;
    const int li = item;
    if (li >= TRACE_ITEMS) return nullptr;
    return reinterpret_cast<long long*>(prm.dbg) + ((long long)(li * NG + g) * MAX_PHASES) * 8;
}

// DBG = false is the production kernel: the layer dump and the timeline stamps are compiled out (the kernel is ~9 k
// instructions; every KB less helps the instruction cache, whose misses show up as stall_no_inst in the epilogues).
template <int MODE, bool DBG, int ACT = ACT_RELU>
__global__ void __launch_bounds__(THREADS, 1) readconv_tc_kernel(const __grid_constant__ TcParams prm) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);     // warp-uniform by construction
    const int lane = threadIdx.x & 31;
    const uint32_t bar0 = ptx::smem_u32(smem + OFF_BAR);
    // barriers: 0,1 w_full[slot]  2,3 w_empty[slot]  4.. act_ready[group]  4+NG.. acc_full[group]  4+2NG.. token[group]
    auto bar = [&](int k) { return bar0 + 8u * k; };
    constexpr int BAR_ACT = 4, BAR_ACC = 4 + NG, BAR_TOK = 4 + 2 * NG, BAR_SFULL = 4 + 3 * NG, BAR_SFREE = 4 + 4 * NG;
    constexpr int W_EPI = NG * EW;                                             // epilogue warps
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

    if (threadIdx.x == 0) {
        ptx::mbar_init(bar(0), 1); ptx::mbar_init(bar(1), 1);
        ptx::mbar_init(bar(2), NG); ptx::mbar_init(bar(3), NG);        // released by every group's issuer
        for (int g = 0; g < NG; ++g) {
            ptx::mbar_init(bar(BAR_ACT + g), EW * 32); ptx::mbar_init(bar(BAR_ACC + g), 1); ptx::mbar_init(bar(BAR_TOK + g), 1);
            ptx::mbar_init(bar(BAR_SFULL + g), 1); ptx::mbar_init(bar(BAR_SFREE + g), EW * 32);
        }
        ptx::fence_mbar_init();
        ptx::mbar_arrive(bar(BAR_TOK));                                // group 0 issues first
    }
    {   // every byte an MMA can read must hold a finite bf16 (zero weights multiply the padding channels)
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (uint32_t i = threadIdx.x; i < OFF_BIAS / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        uint4* zs = reinterpret_cast<uint4*>(smem + OFF_SUM);
        for (uint32_t i = threadIdx.x; i < LOUT * COUT / 4; i += blockDim.x) zs[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (warp == W_EPI + NG) {
        ptx::tmem_alloc(ptx::smem_u32(smem + OFF_TMEM), 512);
        ptx::tmem_relinquish();
    }
    if (threadIdx.x == 32) {
        CtaState* st0 = reinterpret_cast<CtaState*>(smem + OFF_STATE);
        const long long Rt = prm.n_reads;
        auto bound = [&](long long c) -> long long {                       // first read of CTA c
            if (c <= 0) return 0;
            if (c >= (long long)gridDim.x) return Rt;
            if (!prm.allele_out) return min(Rt, (Rt / (NG * G) * c / (long long)gridDim.x) * (NG * G));
            const long long target = Rt * c / (long long)gridDim.x + prm.row_base;
            long long lo = 0, hi = prm.n_alleles;                          // lower bound of `target` in allele_off[0..A]
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if ((long long)__ldg(prm.allele_off + mid) < target) lo = mid + 1; else hi = mid;
            }
            return (long long)__ldg(prm.allele_off + lo) - prm.row_base;
        };
        st0->r_begin = bound(blockIdx.x);
        st0->r_end = bound((long long)blockIdx.x + 1);
        st0->turn = 0;
        if (prm.allele_out && st0->r_begin < st0->r_end) {
            long long lo = 0, hi = prm.n_alleles;                          // allele that starts at r_begin
            const long long target = st0->r_begin + prm.row_base;
            while (lo < hi) {
                const long long mid = (lo + hi) >> 1;
                if ((long long)__ldg(prm.allele_off + mid) < target) lo = mid + 1; else hi = mid;
            }
            st0->cur_allele = (int)lo;
            st0->cur_end = (long long)__ldg(prm.allele_off + lo + 1) - prm.row_base;
        }
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *s_tmem, 0);
    // Work partition: CTA c owns a contiguous range of reads.  In the fused-sum mode the range boundaries are snapped
    // to allele boundaries (lower bound of c*R/grid in the allele CSR), so every allele is summed by one CTA.
    CtaState* cst = reinterpret_cast<CtaState*>(smem + OFF_STATE);
    const long long R = cst->r_end, R0 = cst->r_begin;
    const int n_items_cta = (int)((R - R0 + NG * G - 1) / (NG * G));
    const int n_ph = prm.n_phases;

    if (warp < W_EPI) {
        // ===================================================== epilogue warps (group g = warp / EW)
        const int g = warp / EW, wq = warp & 3, wrow = wq * 32;
        const int tid = threadIdx.x - g * (EW * 32);
        uint8_t* act = smem + OFF_ACT + g * ACT_BYTES;
        const uint32_t tl = tmem_base + ((uint32_t)wrow << 16) + g * GRP_COLS;
        uint32_t acc_n = 0, stg_n = 0;
        float rr[32];                                  // first half of this row's fp32 residual stream
#pragma unroll
        for (int c = 0; c < 32; ++c) rr[c] = 0.f;
        uint32_t raw[IN_IT][2];                        // pileup bytes of the item about to start (8 per row, packed)
        auto fetch_item = [&](int item) {              // item's rows -> `raw` (from the staging buffer when it was staged)
            const long long r0 = R0 + ((long long)item * NG + g) * G;
            const int n = (int)max(0LL, min((long long)G, R - r0));
            if (n <= 0) return;
            const StageDesc sd = stage_desc(prm.reads, prm.n_reads, prm.channels, r0, n);
            if (sd.ok) {
                ptx::mbar_wait(bar(BAR_SFULL + g), stg_n & 1u);            // the producer staged this item's rows
                ++stg_n;
                fetch_rows<true>(raw, smem + OFF_STG + g * STG_BYTES + sd.off, 0, n, prm.channels, prm.layout, tid);
                ptx::mbar_arrive(bar(BAR_SFREE + g));                       // staging may take the next item (the arrive
                                                                            // is ordered after the loads it follows)
            } else {
                fetch_rows<false>(raw, prm.reads, r0, n, prm.channels, prm.layout, tid);
            }
        };
        if (n_items_cta > 0) fetch_item(0);
        for (int item = 0; item < n_items_cta; ++item) {
            const long long r0 = R0 + ((long long)item * NG + g) * G;
            const int n = (int)max(0LL, min((long long)G, R - r0));
            if (n <= 0) continue;
            long long* tr = DBG ? trace_slot(prm, item, g) : nullptr;
            // item-level stamps live in the first spare phase slot: {operand store start, end, after the arrive;
            // issuer: weights landed, operand seen, token received} for phase 0
            if (tr && tid == 0) tr[N_PHASES * 8 + 0] = clock64();
            store_rows(act, raw, tid);
            if (tr && tid == 0) tr[N_PHASES * 8 + 1] = clock64();
            ptx::tc_fence_before();
            ptx::fence_proxy_async();
            ptx::mbar_arrive(bar(BAR_ACT + g));
            if (tr && tid == 0) tr[N_PHASES * 8 + 2] = clock64();
            float* gout = prm.out ? prm.out + r0 * (long long)(LOUT * COUT) : nullptr;
#pragma unroll 1
            for (int ph = 0; ph < n_ph; ++ph) {
                ptx::mbar_wait(bar(BAR_ACC + g), acc_n & 1);
                ++acc_n;
                ptx::tc_fence_after();
                if (tr && tid == 0) tr[ph * 8 + 2] = clock64();
                float* dbg = (DBG && prm.dbg && prm.dbg_phase == ph) ? prm.dbg + (r0 / G) * (T1 * 128 * 64) : nullptr;
                if (ph == 0) {
                    epi_conv<MODE, ACT, true, 16, T1, P1, LV1, false, false, false, false, OUT_NAT, 0>(
                        prm, act, tl, B_L1, 0, n, A1_CH, 2 * A1_CH, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                } else if (ph == 1) {
                    epi_conv<MODE, ACT, true, 16, T1, P1, LV2, false, false, false, false, OUT_Q4, 0>(
                        prm, act, tl, B_L2, 0, n, Q4_ARR, 8 * Q4_ARR, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                } else if (ph == 2) {
                    epi_pool<MODE, ACT>(prm, act, tl, B_L3, n, g, wq, lane, dbg, rr);
                } else if (ph < 9) {
                    const int b = B_S2 + (ph - 3) * 32;
                    // space-to-depth rows (two positions per row): same tile shape and operand layout as the 64-channel stage;
                    // the last layer's output is already the (chunk, parity) layout the stride-2 block reads
                    if ((ph - 3) % 2 == 0)
                        epi_conv<MODE, ACT, true, 64, T3, P3, LV4, false, false, false, false, OUT_NAT, 1, true>(
                            prm, act, tl, b, 0, n, D2_ARR, 8 * D2_ARR, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                    else if (ph < 8)
                        epi_conv<MODE, ACT, true, 64, T3, P3, LV4, true, false, true, false, OUT_NAT, 1, true>(
                            prm, act, tl, b, 0, n, D2_ARR, 8 * D2_ARR, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                    else
                        epi_conv<MODE, ACT, true, 64, T3, P3, LV4, true, false, false, false, OUT_NAT, 1, true>(
                            prm, act, tl, b, 0, n, D2_ARR, 8 * D2_ARR, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                } else if (ph == 9) {
                    epi_conv<MODE, ACT, false, 64, T3, P3, LV4, false, false, false, true, OUT_NAT, 1>(
                        prm, act, tl, B_RCA, 0, n, S3_CH, 8 * S3_CH, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                } else if (ph == 10) {
                    epi_conv<MODE, ACT, true, 64, T3, P3, LV4, true, true, true, false, OUT_NAT, 1>(
                        prm, act, tl, B_RCB, B_RCS, n, S3_CH, 8 * S3_CH, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                } else {
                    const int b = B_S3 + (ph - 11) * 64;
                    if ((ph - 11) % 2 == 0)
                        epi_conv<MODE, ACT, true, 64, T3, P3, LV4, false, false, false, false, OUT_NAT, 1>(
                            prm, act, tl, b, 0, n, S3_CH, 8 * S3_CH, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                    else if (ph + 1 < n_ph)
                        epi_conv<MODE, ACT, true, 64, T3, P3, LV4, true, false, true, false, OUT_NAT, 1>(
                            prm, act, tl, b, 0, n, S3_CH, 8 * S3_CH, nullptr, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                    else {
                        epi_conv<MODE, ACT, true, 64, T3, P3, LV4, true, false, false, false, OUT_GLOBAL, 1>(
                            prm, act, tl, b, 0, n, 0, 0, gout, dbg, wrow, lane, rr, tr ? tr + ph * 8 : nullptr);
                        // the next item's bytes travel to registers while this group waits for its turn in the sum
                        if (item + 1 < n_items_cta) fetch_item(item + 1);
                        if (prm.allele_out) accumulate_out(smem, act, prm, r0, n, item * NG + g, g, tid);
                        else copy_out(act, gout, n, g, tid);
                    }
                }
                if (tr && tid == 0) tr[ph * 8 + 3] = clock64();
                if (ph + 1 < n_ph) {
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::mbar_arrive(bar(BAR_ACT + g));
                }
                if (tr && tid == 0) tr[ph * 8 + 7] = clock64();
            }
        }
    } else if (warp < W_EPI + NG) {
        // ===================================================== MMA issuers: warp W_EPI + g serves group g.
        // The whole warp runs the (warp-uniform) loop; one elected lane issues each tcgen05.mma / commit.
        const int g = warp - W_EPI;
        uint32_t w_n = 0, ar_n = 0, tok_n = 0;
        const uint32_t act_lo = (ptx::smem_u32(smem + OFF_ACT) + g * ACT_BYTES) >> 4;
        const uint32_t w0_lo = ptx::smem_u32(smem + OFF_W) >> 4;
        const uint32_t d0 = tmem_base + g * GRP_COLS;
        for (int item = 0; item < n_items_cta; ++item) {
            const int n = (int)max(0LL, min((long long)G, R - (R0 + ((long long)item * NG + g) * G)));
            long long* tr = DBG ? trace_slot(prm, item, g) : nullptr;
#pragma unroll 1
            for (int ph = 0; ph < n_ph; ++ph) {
                const uint32_t slot = w_n & 1u;
                // Every issuer waits for the slot (even one whose group is empty in this item): the slot is released
                // only when all have passed it, which keeps them in lockstep with the producer.
                ptx::mbar_wait(bar(slot), (w_n >> 1) & 1u);              // this layer's weights have landed
                if (tr && ph == 0 && lane == 0) tr[N_PHASES * 8 + 3] = clock64();
                if (n > 0) {
                    ptx::mbar_wait(bar(BAR_ACT + g), ar_n & 1u);         // this group's operand is written
                    ++ar_n;
                }
                if (tr && ph == 0 && lane == 0) tr[N_PHASES * 8 + 4] = clock64();
                // The groups take turns on the tensor pipe (token passed round-robin): one group's layer executes
                // as a block and its epilogue then overlaps the other group's MMAs.  Without the turn order the
                // issue streams interleave in the pipe's FIFO, all accumulators complete together and all
                // epilogues run together with the tensor pipe idle.
                ptx::mbar_wait(bar(BAR_TOK + g), tok_n & 1u);
                ++tok_n;
                if (tr && ph == 0 && lane == 0) tr[N_PHASES * 8 + 5] = clock64();
                if (n > 0) {
                    ptx::tc_fence_after();
                    if (tr && lane == 0) tr[ph * 8 + 0] = clock64();
                    const uint32_t w_lo = w0_lo + slot * (WSLOT_BYTES >> 4);
                    issue_phase<MODE, 0>(ph, act_lo, w_lo, d0);
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(bar(BAR_TOK + (g + 1) % NG));
                    __syncwarp();
                    issue_phase<MODE, 1>(ph, act_lo, w_lo, d0);
                    ptx::tc_commit(bar(BAR_ACC + g));                    // accumulators ready -> epilogue
                    ptx::tc_commit(bar(2 + slot));                       // weight slot no longer read by this group
                    if (tr && lane == 0) tr[ph * 8 + 1] = clock64();
                } else if (lane == 0) {
                    ptx::mbar_arrive(bar(2 + slot));
                    ptx::mbar_arrive(bar(BAR_TOK + (g + 1) % NG));
                }
                __syncwarp();
                ++w_n;
            }
        }
    } else {
        // ===================================================== weight producer (one thread, bulk async copies)
        if (lane == 0) {
            uint32_t w_n = 0;
            const uint32_t w0 = ptx::smem_u32(smem + OFF_W);
            uint32_t stg_n[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) stg_n[g] = 0;
            auto stage_item = [&](int item) {                      // raw rows of work item `item` -> the groups' staging buffers
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    const long long r0 = R0 + ((long long)item * NG + g) * G;
                    const int n = (int)max(0LL, min((long long)G, R - r0));
                    const StageDesc sd = stage_desc(prm.reads, prm.n_reads, prm.channels, r0, n);
                    if (!sd.ok) continue;
                    if (stg_n[g] > 0) ptx::mbar_wait(bar(BAR_SFREE + g), (stg_n[g] - 1) & 1u);   // previous rows consumed
                    ++stg_n[g];
                    ptx::mbar_expect_tx(bar(BAR_SFULL + g), sd.bytes);
                    ptx::bulk_g2s(ptx::smem_u32(smem + OFF_STG + g * STG_BYTES), sd.src, sd.bytes, bar(BAR_SFULL + g));
                }
            };
            if (n_items_cta > 0) stage_item(0);
            for (int item = 0; item < n_items_cta; ++item) {
                {   // pull the next work item's pileup rows into L2 while this one computes
                    const long long nr0 = R0 + (long long)(item + 1) * (NG * G);
                    if (nr0 < R) {
                        const long long bytes = min((long long)(NG * G), R - nr0) * (LIN * prm.channels);
                        const uint8_t* p = prm.reads + nr0 * (LIN * prm.channels);
                        const uint8_t* p16 = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(p) & ~uintptr_t(15));
                        ptx::prefetch_l2(p16, (uint32_t)(((p - p16) + bytes) & ~15LL));
                    }
                }
#pragma unroll 1
                for (int ph = 0; ph < n_ph; ++ph) {
                    const uint32_t slot = w_n & 1u;
                    ptx::mbar_wait(bar(2 + slot), ((w_n >> 1) & 1u) ^ 1u);
                    const uint32_t bytes = prm.w_bytes[ph];
                    ptx::mbar_expect_tx(bar(slot), bytes);
                    const uint8_t* src = prm.weights + prm.w_src[ph];
                    for (uint32_t o = 0; o < bytes; o += 8192u)
                        ptx::bulk_g2s(w0 + slot * WSLOT_BYTES + o, src + o, min(8192u, bytes - o), bar(slot));
                    ++w_n;
                    // by now every group has built this item's first operand: its staging buffer takes the next item
                    if (ph == 2 && item + 1 < n_items_cta) stage_item(item + 1);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == W_EPI + NG) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host side
inline uint16_t bf16_rne(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline float bf16_to_float(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

struct HostConv {
    const float* w;   // [k*cin][cout]  (row = tap*cin + ci)
    const float* b;
    int cin, cout, k;
};

}  // namespace tc

struct ReadConvTC {
    tc::TcParams prm;
    uint8_t* d_weights = nullptr;
    int mode = 3;
    int act = ACT_RELU;                // the one activation of every convolution in the stack (ReLU or Softplus)
    int sm_count = 148;
};

// Checks that `net` is the read_convolver architecture this kernel is specialised for and packs its weights.
// `d_base` / `h_base`: device and host copies of the same float array the ConvDesc pointers index into.
// `covered` receives the number of leading layer records of `net` the kernel runs (the standard 11, plus up to
// MAX_EXTRA_BLOCKS appended 64-channel residual blocks); the caller runs whatever follows layer by layer.
static ReadConvTC* readconv_tc_create(const std::vector<LayerDesc>& net, const float* d_base, const float* h_base,
                                      int channels, int feature_length, int precision, std::string& err,
                                      size_t* covered = nullptr) {
    using namespace tc;
    if (precision != HELLO_PREC_BF16X3 && precision != HELLO_PREC_BF16) { err = "unknown tensor-core precision"; return nullptr; }
    if (feature_length != LIN || channels < 1 || channels > 8) { err = "tensor-core read convolver needs L=150, C<=8"; return nullptr; }
    // one activation throughout: ReLU, or Softplus (moe_attention_config_single_tech_old_equivalent_layer_norm.py)
    const int act = net.empty() ? ACT_RELU : net[0].a.relu;
    if (act != ACT_RELU && act != ACT_SOFTPLUS) { err = "tensor-core read convolver needs ReLU or Softplus"; return nullptr; }
    auto is_conv = [&](const LayerDesc& L, int cin, int cout, int k, int s, int p) {
        return L.kind == KIND_CONV && L.a.cin == cin && L.a.cout == cout && L.a.k == k && L.a.stride == s && L.a.pad == p && L.a.relu == act;
    };
    auto is_res = [&](const LayerDesc& L, int cin, int cout, int s, bool sc) {
        return L.kind == KIND_RES && L.a.relu == act && L.b.relu == act &&
               L.a.cin == cin && L.a.cout == cout && L.a.k == 3 && L.a.stride == s && L.a.pad == 1 &&
               L.b.cin == cout && L.b.cout == cout && L.b.k == 3 && L.b.stride == 1 && L.b.pad == 1 &&
               (L.has_shortcut != 0) == sc && (!sc || (L.s.cin == cin && L.s.cout == cout && L.s.k == 1 && L.s.stride == s && L.s.pad == 0));
    };
    bool ok = net.size() >= (size_t)N_RECORDS && is_conv(net[0], channels, 16, 3, 1, 0) && is_conv(net[1], 16, 16, 3, 1, 0) &&
              is_conv(net[2], 16, 32, 3, 1, 0) && net[3].kind == KIND_MAXPOOL && net[3].a.k == 3 && net[3].a.stride == 2;
    for (int i = 4; ok && i < 7; ++i) ok = is_res(net[i], 32, 32, 1, false);
    ok = ok && is_res(net[7], 32, 64, 2, true);
    for (int i = 8; ok && i < 11; ++i) ok = is_res(net[i], 64, 64, 1, false);
    if (!ok) { err = "layer table is not the 16-16-32 / 3xRes32 / Res32->64(s2) / 3xRes64 read convolver"; return nullptr; }
    int extra = 0;
    while (extra < MAX_EXTRA_BLOCKS && (size_t)(N_RECORDS + extra) < net.size() && is_res(net[N_RECORDS + extra], 64, 64, 1, false))
        ++extra;
    if (covered) *covered = N_RECORDS + extra;
    else if (net.size() != (size_t)(N_RECORDS + extra)) { err = "layers after the read convolver"; return nullptr; }

    const int parts = precision == HELLO_PREC_BF16X3 ? 2 : 1;
    auto hc = [&](const ConvDesc& c) { return HostConv{h_base + (c.w - d_base), h_base + (c.b - d_base), c.cin, c.cout, c.k}; };
    std::vector<uint8_t> blob;
    std::vector<float> bias(N_BIAS, 0.f);
    ReadConvTC* t = new ReadConvTC();
    t->mode = precision == HELLO_PREC_BF16X3 ? 3 : 1;
    t->act = act;
    std::memset(&t->prm, 0, sizeof(t->prm));

    // One B unit covers (tap, k16 step j): [2 chunks][rows][8] bf16, rows = the cout "hi" rows followed (bf16x3) by the
    // cout "lo" rows; element (chunk c, row n, e) = W[n][ci = 16j+8c+e][tap].
    // `stem1` packs layer 1 instead: unit u covers real taps 2u (chunk 0) and 2u+1 (chunk 1), e = channel (zero padded).
    auto pack_phase = [&](int ph, std::vector<std::pair<HostConv, bool>> convs, int expect_units) {
        std::vector<uint16_t> buf;
        int units = 0, cout = convs[0].first.cout;
        for (auto& cv : convs) {
            const HostConv& c = cv.first;
            const bool stem1 = cv.second;
            if (c.cout != cout) return false;
            const int units_t = stem1 ? 2 : c.k, k16 = stem1 ? 1 : c.cin / 16;
            for (int tp = 0; tp < units_t; ++tp)
                for (int j = 0; j < k16; ++j, ++units)
                    for (int ch = 0; ch < 2; ++ch)
                        for (int part = 0; part < parts; ++part)
                            for (int n = 0; n < c.cout; ++n)
                                for (int e = 0; e < 8; ++e) {
                                    float w = 0.f;
                                    if (stem1) {
                                        const int tap = 2 * tp + ch;
                                        if (tap < c.k && e < c.cin) w = c.w[(size_t)(tap * c.cin + e) * c.cout + n];
                                    } else {
                                        const int ci = 16 * j + 8 * ch + e;
                                        w = c.w[(size_t)(tp * c.cin + ci) * c.cout + n];
                                    }
                                    const uint16_t h = bf16_rne(w);
                                    buf.push_back(part == 0 ? h : bf16_rne(w - bf16_to_float(h)));
                                }
        }
        t->prm.w_src[ph] = (uint32_t)blob.size();
        t->prm.w_bytes[ph] = (uint32_t)(buf.size() * 2);
        const uint8_t* pb = reinterpret_cast<const uint8_t*>(buf.data());
        blob.insert(blob.end(), pb, pb + buf.size() * 2);
        return units == expect_units && t->prm.w_bytes[ph] == (uint32_t)(units * parts * cout * 32) &&
               t->prm.w_bytes[ph] <= WSLOT_BYTES && t->prm.w_bytes[ph] % 16 == 0;
    };
    auto copy_bias = [&](const HostConv& c, int off) { for (int i = 0; i < c.cout; ++i) bias[off + i] = c.b[i]; };
    // A 32 -> 32 k=3 pad=1 convolution in the space-to-depth form of issue_stage2_s2d.  w_t multiplies x[p + t - 1]:
    //   even output 2r:   w0 * odd[r-1]  + w1 * even[r] + w2 * odd[r]
    //   odd output 2r+1:  w0 * even[r]   + w1 * odd[r]  + w2 * even[r+1]
    // Row groups of 32 output channels; element (chunk c, row n, e) = W_t[n][ci = 16 j + 8 c + e] (hi or lo part).
    auto pack_s2d_phase = [&](int ph, const HostConv& c) {
        if (c.cin != 32 || c.cout != 32 || c.k != 3) return false;
        std::vector<uint16_t> buf;
        // groups: (tap, lo?) per 32 rows, in accumulator-column order
        auto unit = [&](int j, std::vector<std::pair<int, bool>> groups) {
            for (int ch = 0; ch < 2; ++ch)
                for (auto& gp : groups)
                    for (int n = 0; n < 32; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int ci = 16 * j + 8 * ch + e;
                            const float w = c.w[(size_t)(gp.first * c.cin + ci) * c.cout + n];
                            const uint16_t h = bf16_rne(w);
                            buf.push_back(gp.second ? bf16_rne(w - bf16_to_float(h)) : h);
                        }
        };
        const bool x3 = parts == 2;
        using G2 = std::vector<std::pair<int, bool>>;
        // centre, even input: even output takes w1, odd output takes w0;  centre, odd input: even <- w2, odd <- w1
        for (int par = 0; par < 2; ++par)
            for (int j = 0; j < 2; ++j) {
                const int te = par == 0 ? 1 : 2, to = par == 0 ? 0 : 1;
                unit(j, x3 ? G2{{te, true}, {te, false}, {to, false}, {to, true}} : G2{{te, false}, {to, false}});
            }
        for (int j = 0; j < 2; ++j) unit(j, x3 ? G2{{0, true}, {0, false}} : G2{{0, false}});      // tap -1: odd[r-1] -> even, w0
        for (int j = 0; j < 2; ++j) unit(j, x3 ? G2{{2, false}, {2, true}} : G2{{2, false}});      // tap +1: even[r+1] -> odd, w2
        t->prm.w_src[ph] = (uint32_t)blob.size();
        t->prm.w_bytes[ph] = (uint32_t)(buf.size() * 2);
        const uint8_t* pb = reinterpret_cast<const uint8_t*>(buf.data());
        blob.insert(blob.end(), pb, pb + buf.size() * 2);
        const uint32_t want = x3 ? UNITS_S2_FULL * s2d_full_bytes<3>() + UNITS_S2_HALF * s2d_half_bytes<3>()
                                 : UNITS_S2_FULL * s2d_full_bytes<1>() + UNITS_S2_HALF * s2d_half_bytes<1>();
        return t->prm.w_bytes[ph] == want && want <= WSLOT_BYTES;
    };

    bool fit = true;
    // stem
    fit &= pack_phase(0, {{hc(net[0].a), true}}, UNITS_L1);
    copy_bias(hc(net[0].a), B_L1);
    fit &= pack_phase(1, {{hc(net[1].a), false}}, UNITS_L2);
    copy_bias(hc(net[1].a), B_L2);
    fit &= pack_phase(2, {{hc(net[2].a), false}}, UNITS_L3);
    copy_bias(hc(net[2].a), B_L3);
    // three residual blocks at 32 channels
    for (int r = 0; r < 3; ++r) {
        const LayerDesc& L = net[4 + r];
        for (int h2 = 0; h2 < 2; ++h2) {
            const int ph = 3 + 2 * r + h2;
            const HostConv c = hc(h2 ? L.b : L.a);
            fit &= pack_s2d_phase(ph, c);
            copy_bias(c, B_S2 + (ph - 3) * 32);
        }
    }
    // stride-2 block 32 -> 64: conv_a and the 1x1 shortcut read the de-interleaved stage-2 output
    fit &= pack_phase(9, {{hc(net[7].a), false}, {hc(net[7].s), false}}, UNITS_RC);
    copy_bias(hc(net[7].a), B_RCA);
    copy_bias(hc(net[7].s), B_RCS);
    fit &= pack_phase(10, {{hc(net[7].b), false}}, UNITS_S3);
    copy_bias(hc(net[7].b), B_RCB);
    t->prm.n_phases = N_PHASES + 2 * extra;
    for (int r = 0; r < 3 + extra; ++r) {
        const LayerDesc& L = net[8 + r];
        for (int h2 = 0; h2 < 2; ++h2) {
            const int ph = 11 + 2 * r + h2;
            const HostConv c = hc(h2 ? L.b : L.a);
            fit &= pack_phase(ph, {{c, false}}, UNITS_S3);
            copy_bias(c, B_S3 + (ph - 11) * 64);
        }
    }
    if (!fit) { err = "a layer's weights do not fit the shared-memory weight slot"; delete t; return nullptr; }

    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { err = "cudaGetDeviceProperties failed"; delete t; return nullptr; }
    t->sm_count = prop.multiProcessorCount;
    if ((size_t)prop.sharedMemPerBlockOptin < SMEM_BYTES) { err = "device has too little shared memory per block"; delete t; return nullptr; }
    if (cudaMalloc(&t->d_weights, blob.size()) != cudaSuccess ||
        cudaMemcpy(t->d_weights, blob.data(), blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        err = "allocating the packed bf16 weights failed";
        if (t->d_weights) cudaFree(t->d_weights);
        delete t;
        return nullptr;
    }
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](const void* fn) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    };
    if (t->mode == 3) {
        opt_in((const void*)readconv_tc_kernel<3, false>); opt_in((const void*)readconv_tc_kernel<3, true>);
        opt_in((const void*)readconv_tc_kernel<3, false, ACT_SOFTPLUS>);
    } else {
        opt_in((const void*)readconv_tc_kernel<1, false>); opt_in((const void*)readconv_tc_kernel<1, true>);
        opt_in((const void*)readconv_tc_kernel<1, false, ACT_SOFTPLUS>);
    }
    if (e != cudaSuccess) {
        err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        cudaFree(t->d_weights); delete t;
        return nullptr;
    }
    t->prm.weights = t->d_weights;
    std::memcpy(t->prm.bias_tab, bias.data(), N_BIAS * sizeof(float));
    t->prm.channels = channels;
    return t;
}

// out: fp32 [n_reads, 36, 64] channel-last.  dbg (optional): fp32 [ceil(n_reads/3), 512, 64] dump of the epilogue
// values of layer phase `dbg_phase` (test hook).
// allele_out / allele_off / n_alleles / row_base (optional): write per-allele sums instead of per-read maps.
static cudaError_t readconv_tc_launch(ReadConvTC* t, const uint8_t* reads, long long n_reads, int layout, float* out,
                                      cudaStream_t st, float* dbg = nullptr, int dbg_phase = -1,
                                      float* allele_out = nullptr, const int32_t* allele_off = nullptr,
                                      long long n_alleles = 0, int row_base = 0) {
    if (n_reads <= 0) return cudaSuccess;
    tc::TcParams prm = t->prm;
    prm.reads = reads;
    prm.n_reads = n_reads;
    prm.layout = layout;
    prm.out = out;
    prm.dbg = dbg;
    prm.dbg_phase = dbg_phase;
    prm.allele_out = allele_out;
    prm.allele_off = allele_off;
    prm.n_alleles = n_alleles;
    prm.row_base = row_base;
    const long long items = (n_reads + tc::NG * tc::G - 1) / (tc::NG * tc::G);
    if (items > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int grid = (int)std::min<long long>(items, t->sm_count);
    const bool debug = dbg != nullptr;
    if (t->act == ACT_SOFTPLUS) {
        if (debug) return cudaErrorNotSupported;       // the layer dump (test hook) exists for the ReLU kernels only
        if (t->mode == 3) tc::readconv_tc_kernel<3, false, ACT_SOFTPLUS><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(prm);
        else tc::readconv_tc_kernel<1, false, ACT_SOFTPLUS><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(prm);
    } else if (t->mode == 3) {
        if (debug) tc::readconv_tc_kernel<3, true><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(prm);
        else tc::readconv_tc_kernel<3, false><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(prm);
    } else {
        if (debug) tc::readconv_tc_kernel<1, true><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(prm);
        else tc::readconv_tc_kernel<1, false><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(prm);
    }
    return cudaGetLastError();
}

static void readconv_tc_destroy(ReadConvTC* t) {
    if (!t) return;
    if (t->d_weights) cudaFree(t->d_weights);
    delete t;
}

}  // namespace hello
