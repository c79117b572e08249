// Generic tcgen05 Conv1d layer (+bias, +ReLU, +residual): the tensor-core counterpart of conv_fp32.cuh for every layer
// that is not inside one of the fused kernels -- meta_convolver_ref (architectures/meta_convolver_ref.py:13-106), the
// layers appended beyond the fused kernels' reach by the addendum models, and all sub-networks of the 2x-wide models
// (architectures/*_wide.py), which do not fit the fused kernels' shared-memory layouts.  Replaces torch.nn.Conv1d called
// through NNTools.WeightNormedConv1d (python/NNTools.py:791-799).
//
// Implicit GEMM without im2col, persistent (two CTAs per SM -- one where a layer's accumulator needs more than half of
// tensor memory -- walk (row tile, column tile) work items):
//   * Rows.  The items (reads / alleles / sites) are laid back to back along M with a fixed PITCH of output rows; the
//     pitch - lout surplus rows of an item are garbage outputs that are never stored.  A 128-row tile of that padded
//     sequence needs the input rows of the same range (+ k - 1), so the tile's input is STAGED ONCE in shared memory in
//     the K-major no-swizzle operand layout (per 8 channels one array of 16-byte rows) and tap t of the convolution is
//     the same array read `delta_t` rows further down -- the A descriptor just starts later (as in readconv_tc.cuh).
//     Rows outside an item are stored as zeros: that is the convolution's padding.  Stride-2 layers read STRIDE arrays
//     (input position modulo the stride), which turns the stride back into unit row shifts:
//         tap t reads position  s*p - pad + t  =  array (t - pad) mod s,  row  p + floor((t - pad) / s) - sigma_min.
//   * K is walked in chunks of KC channels (64, or 32 for stride 2) through a two-slot ring of such staged tiles:
//     eight producer warps (thread = row, half of the chunk's channels each) load fp32 channel-last rows with 128-bit
//     loads, split them into bf16 hi + lo and store the operand; one thread streams the packed weight units of the
//     chunk (tap x 16 channels each) from L2 with cp.async.bulk through a 4-slot ring; one warp issues the MMAs.
//   * bf16x3 is issued in the stacked form where the accumulator fits (Cout tile <= 128): A_hi x [W_hi | W_lo] and
//     A_lo x W_hi -- two instructions per (tap, 16 channels), the 4 KB activation operand fetched twice instead of three
//     times (tools/mma_bench.cu: an MMA with both operands in shared memory costs (4096 + 32 N) / 128 cycles); a
//     256-column tile takes three instructions into one accumulator.
//   * Two accumulators in tensor memory (2 x 256 columns): four epilogue warps (TMEM -> bias / ReLU / residual -> HBM)
//     work on tile i while the MMAs of tile i + 1 run.
// Layer by layer through HBM (fp32 activations): these layers are bound by that traffic, not by the tensor pipe.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/hello_moe.h"
#include "common.cuh"
#include "tc_ptx.cuh"
#include "readconv_tc.cuh"   // bf16 host helpers, store_chunk8

namespace hello {
namespace cl {

constexpr int E_WARPS = 4, P_WARPS = 8, P_GROUP = 128;  // two producer groups of four warps take turns on the K chunks
constexpr int THREADS = (E_WARPS + P_WARPS + 2) * 32;     // + MMA issuer + weight loader
constexpr int B_STAGES = 4, A_SLOTS = 2;
constexpr int ROWS_TAB = 136;                             // rows staged per input array: 128 + largest tap shift, rounded up to 8
// bytes of one 8-channel operand array: one row more than staged, so that consecutive arrays start 16 bytes apart modulo
// 128 and the producers' 8-byte stores (lanes = 8-channel halves of one row) spread over all banks
constexpr uint32_t ARR = (ROWS_TAB + 1) * 16;
constexpr int MAX_TAPS = 3;
constexpr int EPI_COLS = 64;                              // columns per epilogue block (one 256-byte piece of a row)
constexpr uint32_t STG_PITCH = EPI_COLS * 4 + 16;         // row pitch of the epilogue staging tiles: 16-byte skew per row
// every CTA takes all 512 tensor-memory columns: ask for more than half an SM's shared memory so that two never share one
constexpr size_t MIN_SMEM = 117 * 1024;

struct ConvTcArgs {
    const float* x;          // fp32 channel-last input [n_items][lin][cin] (item stride sn, row stride sl floats)
    long long sn, sl;
    const uint8_t* w;        // packed units: [column tile][K chunk][tap][16-channel step]
    const float* bias;
    float* y;                // [n_items * lout][cout]
    const float* resid;      // same shape or nullptr, added after the ReLU
    uint32_t n_items, total_rows;          // total_rows = n_items * pitch (padded output rows)
    uint32_t m_tiles, n_tiles;
    int lin, lout, cin, cout, ksz, stride, relu, nt, stacked;
    int pitch, sigma_min, n_arr, kc, n_chunks, max_delta;
    int tap_arr[MAX_TAPS], tap_delta[MAX_TAPS];
    long long* trace;        // developer timeline (tools/convlayer_trace.cu): CTA 0 stamps clock64() per tile and role; else nullptr
    int acc_shift;           // log2 of the accumulators in tensor memory (1: two, tile i+1 fills one while tile i is read; 0: one)
    int res_ring;            // two CTAs per SM: the residual comes through a two-slot ring of 16-column pieces per row
};

// Timeline slots of one tile: producer {slot free, loads issued, operand stored}, issuer {accumulator free, operand
// seen, last MMA issued}, epilogue {accumulator seen, rows written}.
constexpr int TRACE_TILES = 48, TRACE_SLOTS = 8;
__device__ __forceinline__ void trace_stamp(const ConvTcArgs& a, uint32_t tile_seq, int slot) {
    if (a.trace && blockIdx.x == 0 && tile_seq < (uint32_t)TRACE_TILES) a.trace[tile_seq * TRACE_SLOTS + slot] = clock64();
}

// shared -> global bulk copy of the issuing thread's own bulk group (TMA unit); the source must stay untouched until
// bulk_wait_read
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 16-byte asynchronous copy global -> shared (LDGSTS, past L1), in the issuing thread's own groups
__device__ __forceinline__ void ldgsts16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ldgsts_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ldgsts_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
constexpr uint32_t RING_PITCH = 2 * 64 + 16;              // residual ring of one row: two 16-column slots, 16-byte skew per row

// CPS = CTAs per SM.  2: half of tensor memory (256 columns, accumulators 128 columns apart), at most half of the shared
// memory (the residual comes through a small ring instead of a staging tile) and 72 registers, so that two CTAs
// share an SM: every warp role here is a few warps running dependent chains, and a second independent pipeline on the SM
// fills the issue slots, the tensor pipe and the memory pipes the first one leaves idle.
template <int MODE, int CPS>
__global__ void __launch_bounds__(THREADS, CPS) convlayer_tc_kernel(const __grid_constant__ ConvTcArgs a) {
    constexpr uint32_t TMEM_COLS = CPS == 2 ? 256u : 512u, ACC_STRIDE = CPS == 2 ? 128u : 256u;
    constexpr int LOAD_BATCH = CPS == 2 ? 9 : 16;         // 128-bit loads a producer thread keeps in flight
    constexpr bool RES_STAGED = CPS == 1;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    constexpr uint32_t PARTS = MODE == 3 ? 2u : 1u;
    const uint32_t c8n = (uint32_t)a.kc / 8;                                  // 8-channel arrays per (part, input array)
    const uint32_t a_part = (uint32_t)a.n_arr * c8n * ARR;                    // hi plane of a staged chunk; lo follows
    const uint32_t a_slot = PARTS * a_part;
    const uint32_t unit = PARTS * 32u * (uint32_t)a.nt;
    uint8_t* s_a = smem;
    uint8_t* s_b = s_a + A_SLOTS * a_slot;
    uint8_t* s_out = s_b + B_STAGES * unit;                                   // epilogue staging: output rows, residual rows
    uint8_t* s_res = s_out + 128 * STG_PITCH;                                // (absent with two CTAs per SM)
    long long* s_tab = reinterpret_cast<long long*>(                          // [2 groups][n_arr][ROWS_TAB]
        s_res + (RES_STAGED ? 128 * STG_PITCH : a.res_ring ? 128 * RING_PITCH : 0));
    const uint32_t bar0 = ptx::smem_u32(s_tab + 2 * 2 * ROWS_TAB);
    auto bar = [&](int k) { return bar0 + 8u * k; };
    constexpr int BAR_AFULL = 0, BAR_AEMPTY = A_SLOTS, BAR_BFULL = 2 * A_SLOTS, BAR_BEMPTY = BAR_BFULL + B_STAGES,
                  BAR_ACCFULL = BAR_BEMPTY + B_STAGES, BAR_ACCEMPTY = BAR_ACCFULL + 2, BAR_RES = BAR_ACCEMPTY + 2,
                  N_BARS = BAR_RES + E_WARPS * 32;
    volatile uint32_t* s_tmem = reinterpret_cast<volatile uint32_t*>(s_tab + 2 * 2 * ROWS_TAB + N_BARS);

    if (threadIdx.x == 0) {
        for (int s = 0; s < A_SLOTS; ++s) { ptx::mbar_init(bar(BAR_AFULL + s), P_GROUP); ptx::mbar_init(bar(BAR_AEMPTY + s), 1); }
        for (int s = 0; s < B_STAGES; ++s) { ptx::mbar_init(bar(BAR_BFULL + s), 1); ptx::mbar_init(bar(BAR_BEMPTY + s), 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(bar(BAR_ACCFULL + s), 1); ptx::mbar_init(bar(BAR_ACCEMPTY + s), E_WARPS * 32); }
        for (int s = 0; s < E_WARPS * 32; ++s) ptx::mbar_init(bar(BAR_RES + s), 1);   // one per epilogue thread: its residual row
        ptx::fence_mbar_init();
    }
    if (warp == E_WARPS + P_WARPS) {
        ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(s_tmem)), TMEM_COLS);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t n_work = a.m_tiles * a.n_tiles;                           // work item = (row tile, column tile), column fastest
    const int k16c = a.kc / 16;                                              // 16-channel steps per chunk

    if (warp < E_WARPS) {
        // ------------------------------------------------------------------ epilogue: thread = row of the tile
        // Every thread moves its own row: the residual row comes in by a bulk copy into the thread's line of a staging tile
        // (prefetched one block ahead, completion on the thread's own mbarrier), the finished row goes out by a bulk
        // copy from the thread's line of a second staging tile -- whole 256-byte pieces of a row on the wire instead of
        // 16-byte pieces from 32 different rows per instruction, and no barrier between the epilogue threads.
        const int r = threadIdx.x;
        const int n_cb = (a.nt + EPI_COLS - 1) / EPI_COLS, cb_cols = a.nt < EPI_COLS ? a.nt : EPI_COLS;
        const uint32_t cb_bytes = (uint32_t)cb_cols * 4u;
        uint8_t* my_out = s_out + (uint32_t)r * STG_PITCH;
        uint8_t* my_res = s_res + (uint32_t)r * STG_PITCH;
        const uint32_t my_bar = bar(BAR_RES + r);
        uint32_t res_n = 0;                                                  // residual pieces consumed so far
        // (row, column block) visited in order: the residual piece of the next one is requested as soon as the staging
        // line is free
        auto row_of = [&](uint32_t wk, bool* ok, long long* orow, int* n0) {
            const uint32_t mt = wk / a.n_tiles, ntile = wk - mt * a.n_tiles;
            const uint32_t g = mt * 128u + (uint32_t)r;
            const uint32_t item = g / (uint32_t)a.pitch, p = g - item * (uint32_t)a.pitch;
            *ok = item < a.n_items && p < (uint32_t)a.lout;
            *orow = (long long)item * a.lout + p;
            *n0 = (int)ntile * a.nt;
        };
        auto request_res = [&](uint32_t wk, int cb) {
            if (!RES_STAGED || !a.resid || wk >= n_work) return;
            bool ok; long long orow; int n0;
            row_of(wk, &ok, &orow, &n0);
            if (!ok) return;
            ptx::mbar_expect_tx(my_bar, cb_bytes);
            ptx::bulk_g2s(ptx::smem_u32(my_res), a.resid + orow * a.cout + n0 + cb * EPI_COLS, cb_bytes, my_bar);
        };
        request_res(blockIdx.x, 0);
        // Two CTAs per SM have no room for a residual tile.  Where a ring of two 16-column pieces per row fits, the residual
        // comes through it two pieces ahead of the arithmetic: 16-byte asynchronous copies, four lanes along the 64 bytes of
        // a row so that a warp instruction fetches whole sectors of eight rows (a thread fetching its own row would touch 32
        // rows and use half of every sector), one group per piece, empty past the end so that "all but the newest group"
        // always means "this piece has landed"; a warp barrier either side of the reads, since the lines a thread reads were
        // filled by other lanes of its warp.  Else the residual is read where it is used.
        const bool ring = !RES_STAGED && a.res_ring;
        const uint32_t my_ring = ptx::smem_u32(s_res + (uint32_t)r * RING_PITCH);
        const uint32_t warp_ring = ptx::smem_u32(s_res + (uint32_t)(warp * 32) * RING_PITCH);
        const int n_j = a.nt / 16;
        uint32_t rq_wk = blockIdx.x, rq_n = 0, rd_n = 0;
        int rq_j = 0;
        auto ring_request = [&]() {
            if (rq_wk < n_work) {
                bool ok; long long orow; int n0;
                row_of(rq_wk, &ok, &orow, &n0);
                const float* src = a.resid + n0 + rq_j * 16 + (lane & 3) * 4;
                const uint32_t dst = warp_ring + (rq_n & 1u) * 64u + (uint32_t)(lane & 3) * 16u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {                                // rows 8q .. 8q+7 of this warp's 32
                    const int row = 8 * q + (lane >> 2);
                    const bool rok = __shfl_sync(0xffffffffu, (int)ok, row);
                    const long long rrow = __shfl_sync(0xffffffffu, orow, row);
                    if (rok) ldgsts16(dst + (uint32_t)row * RING_PITCH, src + rrow * a.cout);
                }
                if (++rq_j == n_j) { rq_j = 0; rq_wk += gridDim.x; }
            }
            ldgsts_commit();
            ++rq_n;
        };
        if (ring) { ring_request(); ring_request(); }
        uint32_t it = 0;
        for (uint32_t wk = blockIdx.x; wk < n_work; wk += gridDim.x, ++it) {
            bool row_ok; long long orow; int n0;
            row_of(wk, &row_ok, &orow, &n0);
            float* yrow = a.y + orow * a.cout + n0;
            const float* rrow = a.resid ? a.resid + orow * a.cout + n0 : nullptr;
            const uint32_t ab = it & (uint32_t)a.acc_shift;
            ptx::mbar_wait(bar(BAR_ACCFULL + ab), (it >> a.acc_shift) & 1u);
            ptx::tc_fence_after();
            if (r == 0) trace_stamp(a, it, 6);
            const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16) + ab * ACC_STRIDE;
            for (int cb = 0; cb < n_cb; ++cb) {
                if (row_ok) bulk_wait_read();                                // the previous piece has left the output line
                if (RES_STAGED && a.resid && row_ok) { ptx::mbar_wait(my_bar, res_n & 1u); ++res_n; }
                __syncwarp();
                for (int cc = 0; cc < cb_cols; cc += 16) {
                    const int c0 = cb * EPI_COLS + cc;
                    float acc[16];
                    ptx::tmem_ld16(tl + c0, acc);
                    if (MODE == 3 && a.stacked) {
                        float hl[16];
                        ptx::tmem_ld16(tl + a.nt + c0, hl);
                        ptx::tmem_wait_ld();
#pragma unroll
                        for (int q = 0; q < 16; ++q) acc[q] += hl[q];
                    } else {
                        ptx::tmem_wait_ld();
                    }
                    if (ring) { ldgsts_wait_but_one(); __syncwarp(); }
                    if (row_ok) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + c0) + q);
                            float4 o = make_float4(acc[4 * q] + b.x, acc[4 * q + 1] + b.y, acc[4 * q + 2] + b.z, acc[4 * q + 3] + b.w);
                            if (a.relu == ACT_RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                            else if (a.relu) { o.x = apply_activation(o.x, a.relu); o.y = apply_activation(o.y, a.relu);
                                               o.z = apply_activation(o.z, a.relu); o.w = apply_activation(o.w, a.relu); }
                            if (a.resid) {
                                float4 t;
                                if (RES_STAGED) t = *reinterpret_cast<const float4*>(my_res + (cc + 4 * q) * 4);
                                else if (ring) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                                            : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                                            : "r"(my_ring + (rd_n & 1u) * 64u + 16u * q) : "memory");
                                else t = __ldg(reinterpret_cast<const float4*>(rrow + c0) + q);
                                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                            }
                            *reinterpret_cast<float4*>(my_out + (cc + 4 * q) * 4) = o;
                        }
                    }
                    if (ring) { __syncwarp(); ring_request(); ++rd_n; }      // this piece's slot is free: fetch the piece after next
                }
                ptx::fence_proxy_async();        // this thread's staging reads / writes are ordered before the bulk copies below
                // the residual line is free: fetch the piece of the next (tile, block)
                if (cb + 1 < n_cb) request_res(wk, cb + 1); else request_res(wk + gridDim.x, 0);
                if (row_ok) bulk_s2g(yrow + cb * EPI_COLS, ptx::smem_u32(my_out), cb_bytes);
                __syncwarp();
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(bar(BAR_ACCEMPTY + ab));                        // the accumulator may be overwritten
            if (r == 0) trace_stamp(a, it, 7);
        }
        bulk_wait_all();
    } else if (warp < E_WARPS + P_WARPS) {
        // ------------------------------------------------------------------ A producers
        // Group pg (four warps) stages the chunks whose running number is pg modulo 2, so one group's loads are in flight
        // while the other converts.  Lanes run along the channels of a row: a warp instruction reads 512 contiguous bytes.
        const int t = threadIdx.x - E_WARPS * 32;
        const int r = t & (P_GROUP - 1), pg = t >> 7;
        long long* tab = s_tab + pg * 2 * ROWS_TAB;
        const int f4_shift = a.kc == 64 ? 4 : a.kc == 32 ? 3 : 2;            // float4 per staged row = kc / 4
        const int n_f4 = (a.n_arr * ROWS_TAB) << f4_shift;
        uint32_t chunk_n = 0, tab_mt = 0xffffffffu;
        for (uint32_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
            const uint32_t mt = wk / a.n_tiles;
            for (int ch = 0; ch < a.n_chunks; ++ch, ++chunk_n) {
                if ((chunk_n & 1u) != (uint32_t)pg) continue;
                if (tab_mt != mt) {
                    // where every staged row of this row tile comes from (element offset into x, -1 = a row of zeros)
                    tab_mt = mt;
                    if (pg == 0) ptx::named_bar_sync(1, P_GROUP); else ptx::named_bar_sync(2, P_GROUP);                    // the previous tile's table is no longer read
                    for (int i = r; i < a.n_arr * ROWS_TAB; i += P_GROUP) {
                        const int arr = i >= ROWS_TAB ? 1 : 0, rr = i - arr * ROWS_TAB;
                        const uint32_t g = mt * 128u + (uint32_t)rr;
                        const uint32_t item = g / (uint32_t)a.pitch;
                        const int q = a.stride * ((int)(g - item * (uint32_t)a.pitch) + a.sigma_min) + arr;
                        const bool ok = rr < 128 + a.max_delta && item < a.n_items && q >= 0 && q < a.lin;
                        tab[i] = ok ? (long long)item * a.sn + (long long)q * a.sl : -1ll;
                    }
                    if (pg == 0) ptx::named_bar_sync(1, P_GROUP); else ptx::named_bar_sync(2, P_GROUP);
                }
                const uint32_t slot = chunk_n % A_SLOTS;
                ptx::mbar_wait(bar(BAR_AEMPTY + slot), ((chunk_n / A_SLOTS) & 1u) ^ 1u);
                if (r == 0 && ch == 0) trace_stamp(a, chunk_n / (uint32_t)a.n_chunks, 0);
                uint8_t* buf = s_a + slot * a_slot;
                const float* xc = a.x + ch * a.kc;
                for (int base = 0; base < n_f4; base += P_GROUP * LOAD_BATCH) {
                    float4 v[LOAD_BATCH];
#pragma unroll
                    for (int u = 0; u < LOAD_BATCH; ++u) {                   // all loads in flight before the first conversion
                        const int i = base + u * P_GROUP + r;
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (i < n_f4) {
                            const long long off = tab[i >> f4_shift];
                            if (off >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(xc + off) + (i & ((1 << f4_shift) - 1)));
                        }
                    }
                    if (r == 0 && ch == 0 && base == 0) trace_stamp(a, chunk_n / (uint32_t)a.n_chunks, 1);
#pragma unroll
                    for (int u = 0; u < LOAD_BATCH; ++u) {
                        const int i = base + u * P_GROUP + r;
                        if (i < n_f4) {
                            const int row = i >> f4_shift, c4 = i & ((1 << f4_shift) - 1);
                            const int arr = row >= ROWS_TAB ? 1 : 0, rr = row - arr * ROWS_TAB;
                            uint8_t* dst = buf + ((uint32_t)arr * c8n + (uint32_t)(c4 >> 1)) * ARR + (uint32_t)rr * 16 + (c4 & 1) * 8;
                            const uint32_t h0 = ptx::pack_bf16x2(v[u].x, v[u].y), h1 = ptx::pack_bf16x2(v[u].z, v[u].w);
                            *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
                            if (MODE == 3) {
                                float d0 = v[u].x, d1 = v[u].y, d2 = v[u].z, d3 = v[u].w;
                                ptx::sub2(d0, d1, __uint_as_float(h0 << 16), __uint_as_float(h0 & 0xffff0000u));
                                ptx::sub2(d2, d3, __uint_as_float(h1 << 16), __uint_as_float(h1 & 0xffff0000u));
                                *reinterpret_cast<uint2*>(dst + a_part) = make_uint2(ptx::pack_bf16x2(d0, d1), ptx::pack_bf16x2(d2, d3));
                            }
                        }
                    }
                }
                ptx::fence_proxy_async();
                ptx::mbar_arrive(bar(BAR_AFULL + slot));
                if (r == 0 && ch == a.n_chunks - 1) trace_stamp(a, chunk_n / (uint32_t)a.n_chunks, 2);
            }
        }
    } else if (warp == E_WARPS + P_WARPS) {
        // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
        const uint32_t idesc_n = ptx::idesc_bf16_m128((uint32_t)a.nt), idesc_2n = ptx::idesc_bf16_m128(2u * (uint32_t)a.nt);
        const uint32_t b_lbo = (MODE == 3 && a.stacked ? 2u : 1u) * (uint32_t)a.nt * 16u;
        uint32_t chunk_n = 0, unit_n = 0, it = 0;
        for (uint32_t wk = blockIdx.x; wk < n_work; wk += gridDim.x, ++it) {
            const uint32_t ab = it & (uint32_t)a.acc_shift;
            ptx::mbar_wait(bar(BAR_ACCEMPTY + ab), ((it >> a.acc_shift) & 1u) ^ 1u);   // the epilogue has read this accumulator
            ptx::tc_fence_after();
            if (lane == 0) trace_stamp(a, it, 3);
            const uint32_t d = tmem + ab * ACC_STRIDE;
            uint32_t first = 1u;
            for (int ch = 0; ch < a.n_chunks; ++ch, ++chunk_n) {
                const uint32_t slot = chunk_n % A_SLOTS;
                ptx::mbar_wait(bar(BAR_AFULL + slot), (chunk_n / A_SLOTS) & 1u);
                ptx::tc_fence_after();
                if (lane == 0 && ch == 0) trace_stamp(a, it, 4);
                const uint32_t abase = ptx::smem_u32(s_a + slot * a_slot);
                for (int tp = 0; tp < a.ksz; ++tp) {
                    for (int j = 0; j < k16c; ++j, ++unit_n) {
                        const uint32_t stage = unit_n % B_STAGES;
                        ptx::mbar_wait(bar(BAR_BFULL + stage), (unit_n / B_STAGES) & 1u);
                        ptx::tc_fence_after();
                        const uint32_t aaddr = abase + ((uint32_t)a.tap_arr[tp] * c8n + 2u * (uint32_t)j) * ARR + (uint32_t)a.tap_delta[tp] * 16u;
                        const uint32_t al = ptx::desc_lo(aaddr, ARR);
                        const uint32_t bl = ptx::desc_lo(ptx::smem_u32(s_b + stage * unit), b_lbo);
                        if (MODE != 3) {
                            ptx::mma_bf16_ss(d, al, bl, idesc_n, first ^ 1u);
                        } else if (a.stacked) {
                            ptx::mma_bf16_ss(d, al, bl, idesc_2n, first ^ 1u);                       // A_hi x [W_hi | W_lo]
                            ptx::mma_bf16_ss(d, al + (a_part >> 4), bl, idesc_n, 1u);                // A_lo x W_hi
                        } else {
                            ptx::mma_bf16_ss(d, al + (a_part >> 4), bl, idesc_n, first ^ 1u);        // A_lo x W_hi
                            ptx::mma_bf16_ss(d, al, bl + ((32u * (uint32_t)a.nt) >> 4), idesc_n, 1u); // A_hi x W_lo
                            ptx::mma_bf16_ss(d, al, bl, idesc_n, 1u);                                // A_hi x W_hi
                        }
                        first = 0u;
                        ptx::tc_commit(bar(BAR_BEMPTY + stage));
                        __syncwarp();
                    }
                }
                ptx::tc_commit(bar(BAR_AEMPTY + slot));
                __syncwarp();
            }
            ptx::tc_commit(bar(BAR_ACCFULL + ab));
            __syncwarp();
            if (lane == 0) trace_stamp(a, it, 5);
        }
    } else {
        // ------------------------------------------------------------------ weight loader
        if (lane == 0) {
            const uint32_t units = (uint32_t)a.ksz * (uint32_t)(a.cin / 16);
            uint32_t unit_n = 0;
            for (uint32_t wk = blockIdx.x; wk < n_work; wk += gridDim.x) {
                const uint32_t ntile = wk % a.n_tiles;
                const uint8_t* src = a.w + (size_t)ntile * units * unit;
                for (uint32_t s = 0; s < units; ++s, ++unit_n) {
                    const uint32_t stage = unit_n % B_STAGES;
                    ptx::mbar_wait(bar(BAR_BEMPTY + stage), ((unit_n / B_STAGES) & 1u) ^ 1u);
                    ptx::mbar_expect_tx(bar(BAR_BFULL + stage), unit);
                    for (uint32_t o = 0; o < unit; o += 8192u)
                        ptx::bulk_g2s(ptx::smem_u32(s_b + stage * unit) + o, src + (size_t)s * unit + o, min(8192u, unit - o),
                                      bar(BAR_BFULL + stage));
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == E_WARPS + P_WARPS) ptx::tmem_dealloc(tmem, TMEM_COLS);
}

// How a layer maps onto the kernel: chunking of K, tap -> (input array, row shift), pitch of an item's output rows.
struct Geometry {
    int nt, stacked, kc, n_arr, stride, sigma_min, max_delta;
    int cps, acc_shift;                                  // CTAs per SM the layer runs with; log2(accumulators in tensor memory)
    int tap_arr[MAX_TAPS], tap_delta[MAX_TAPS];
    int pitch(int lin, int lout) const {
        // rows [0, pitch) of an item are its own; a tap reaching row `pitch + j` reads row j of the next item, which is
        // fine when that row is zero there (a leading padding row) and the reaching tap wants a position beyond lin.
        int p = lout + max_delta;
        bool share = sigma_min < 0 && max_delta >= 1;
        for (int t = 0; t < MAX_TAPS && share; ++t)
            if (tap_delta[t] == max_delta) {
                const int q = stride * (lout - 1 + max_delta + sigma_min) + tap_arr[t];  // position wanted at the last output row
                share = q >= lin;
            }
        return share ? p - 1 : p;
    }
};

// a two-CTA layer with a residual: room for the ring?
inline size_t ring_bytes() { return 128 * RING_PITCH; }
constexpr size_t HALF_SM_SMEM = 113 * 1024;
inline size_t smem_bytes(int mode, const Geometry& g) {
    const size_t parts = mode == 3 ? 2 : 1;
    const size_t a_slot = parts * g.n_arr * (g.kc / 8) * ARR, unit = parts * 32u * g.nt;
    return A_SLOTS * a_slot + B_STAGES * unit + (g.cps == 2 ? 1 : 2) * 128 * STG_PITCH + 2 * 2 * ROWS_TAB * 8 +
           (2 * A_SLOTS + 2 * B_STAGES + 4 + E_WARPS * 32) * 8 + 16;
}

// cps = 2: the layer as two CTAs per SM (half the tensor memory, K chunks of at most 32 channels -- kc_cap = 16 where the
// residual ring needs the room --, no residual staging tile), when its accumulator fits 256 columns and the CTA fits half
// an SM's shared memory
inline bool geometry(const ConvDesc& c, int mode, Geometry* g, int cps = 1, int kc_cap = 64) {
    if (c.k > MAX_TAPS || c.stride < 1 || c.stride > 2 || c.pad < 0) return false;
    g->cps = cps;
    g->nt = c.cout <= 256 ? c.cout : 256;
    g->stacked = mode == 3 && g->nt <= (cps == 2 ? 64 : 128);
    const int acc_cols = g->stacked ? 2 * g->nt : g->nt, acc_stride = cps == 2 ? 128 : 256;
    g->acc_shift = acc_cols <= acc_stride ? 1 : 0;
    g->stride = c.stride;
    auto fdiv = [](int x, int s) { return x >= 0 ? x / s : -((-x + s - 1) / s); };
    int smin = 1 << 30, smax = -(1 << 30);
    for (int t = 0; t < c.k; ++t) { const int s = fdiv(t - c.pad, c.stride); smin = std::min(smin, s); smax = std::max(smax, s); }
    g->sigma_min = smin;
    g->max_delta = smax - smin;
    if (g->max_delta > ROWS_TAB - 128) return false;
    for (int t = 0; t < MAX_TAPS; ++t) { g->tap_arr[t] = 0; g->tap_delta[t] = 0; }
    g->n_arr = 1;                                       // input arrays (position modulo the stride) some tap reads
    for (int t = 0; t < c.k; ++t) {
        const int s = fdiv(t - c.pad, c.stride);
        g->tap_arr[t] = (t - c.pad) - s * c.stride;
        g->tap_delta[t] = s - smin;
        g->n_arr = std::max(g->n_arr, g->tap_arr[t] + 1);
    }
    g->kc = std::min(std::min(c.cin, kc_cap), (cps == 2 ? 32 : 64) / g->n_arr);
    if (c.cin % g->kc || g->kc % 16) return false;
    return cps == 1 || smem_bytes(mode, *g) <= HALF_SM_SMEM;
}

struct PackedConv {
    uint8_t* d_w = nullptr;
    Geometry g;
};

inline bool eligible(const ConvDesc& c) {
    Geometry g;
    return c.cin % 16 == 0 && c.cout % 16 == 0 && c.cin >= 16 && c.cout >= 16 && (c.k == 1 || c.k == 3) &&
           (c.cout <= 256 || c.cout % 256 == 0) && (c.cout <= EPI_COLS || c.cout % EPI_COLS == 0) &&   // whole epilogue blocks
           geometry(c, 3, &g);
}

}  // namespace cl

// Packed weights of every generic tensor-core layer of a handle, keyed by the layer's (device) weight pointer.
struct ConvLayerTC {
    std::map<const float*, cl::PackedConv> layers;
    int mode = 3;
    int sm_count = 148;
    long long* d_trace = nullptr;          // developer timeline buffer (tools/convlayer_trace.cu), nullptr in the library
    int cps = 2;                           // CTAs per SM asked for (2 wherever a layer's geometry allows it)
};

// `with_resid`: the layer's launches add a residual.  Two CTAs per SM then need room for the residual ring; without it the
// rows would be read where they are used, 16 bytes per thread from 32 different rows per instruction, which costs more than
// the second CTA brings (128 -> 128 channels, 1.39 M rows: 0.85 ms against 0.80 ms with one CTA per SM and its staged rows).
// Where the ring does not fit beside K chunks of 32 channels (column tile 128), chunks of 16 make the room: 0.63 ms.
static bool convlayer_tc_add(ConvLayerTC* t, const ConvDesc& c, const float* d_base, const float* h_base, std::string& err,
                             bool with_resid = false) {
    if (!cl::eligible(c) || t->layers.count(c.w)) return true;
    cl::PackedConv p;
    auto fits = [&]() { return !with_resid || cl::smem_bytes(t->mode, p.g) + cl::ring_bytes() <= cl::HALF_SM_SMEM; };
    const bool two = t->cps == 2 && ((cl::geometry(c, t->mode, &p.g, 2) && fits()) ||
                                     (cl::geometry(c, t->mode, &p.g, 2, 16) && fits()));
    if (!two && !cl::geometry(c, t->mode, &p.g, 1)) return true;
    const int parts = t->mode == 3 ? 2 : 1;
    const float* w = h_base + (c.w - d_base);                       // [k*cin][cout]
    const int nt = p.g.nt, kc = p.g.kc;
    // unit (tap, 16 channels) = two 8-channel chunks of `parts * nt` rows of 8 bf16.  Stacked: per chunk the nt "hi" rows
    // followed by the nt "lo" rows (one B operand of 2 nt rows);  otherwise the whole hi unit, then the whole lo unit.
    std::vector<uint16_t> blob;
    auto put = [&](int n0, int tap, int ci0, int ch, int part) {
        for (int n = 0; n < nt; ++n)
            for (int e = 0; e < 8; ++e) {
                const float v = w[(size_t)(tap * c.cin + ci0 + 8 * ch + e) * c.cout + n0 + n];
                const uint16_t h = tc::bf16_rne(v);
                blob.push_back(part == 0 ? h : tc::bf16_rne(v - tc::bf16_to_float(h)));
            }
    };
    for (int n0 = 0; n0 < c.cout; n0 += nt)
        for (int chunk = 0; chunk < c.cin / kc; ++chunk)
            for (int tap = 0; tap < c.k; ++tap)
                for (int j = 0; j < kc / 16; ++j) {
                    const int ci0 = chunk * kc + 16 * j;
                    if (p.g.stacked) {
                        for (int ch = 0; ch < 2; ++ch)
                            for (int part = 0; part < parts; ++part) put(n0, tap, ci0, ch, part);
                    } else {
                        for (int part = 0; part < parts; ++part)
                            for (int ch = 0; ch < 2; ++ch) put(n0, tap, ci0, ch, part);
                    }
                }
    if (cudaMalloc(&p.d_w, blob.size() * 2) != cudaSuccess ||
        cudaMemcpy(p.d_w, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess) {
        err = "allocating packed layer weights failed";
        if (p.d_w) cudaFree(p.d_w);
        return false;
    }
    t->layers[c.w] = p;
    return true;
}

static size_t convlayer_tc_smem(int mode, const cl::Geometry& g) { return cl::smem_bytes(mode, g); }

static ConvLayerTC* convlayer_tc_create(int precision, std::string& err) {
    ConvLayerTC* t = new ConvLayerTC();
    t->mode = precision == HELLO_PREC_BF16X3 ? 3 : 1;
    const int smem = 227 * 1024;             // opt in to everything; a launch asks for what its geometry needs
    cudaError_t e = cudaSuccess;
    auto opt_in = [&](const void* fn) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    };
    if (t->mode == 3) { opt_in((const void*)cl::convlayer_tc_kernel<3, 1>); opt_in((const void*)cl::convlayer_tc_kernel<3, 2>); }
    else { opt_in((const void*)cl::convlayer_tc_kernel<1, 1>); opt_in((const void*)cl::convlayer_tc_kernel<1, 2>); }
    if (const char* env = std::getenv("HELLO_CL_CPS")) t->cps = std::atoi(env) == 1 ? 1 : 2;      // developer A/B switch
    cudaDeviceProp prop;
    int dev = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { err = std::string("convlayer_tc_create: ") + cudaGetErrorString(e); delete t; return nullptr; }
    t->sm_count = prop.multiProcessorCount;
    return t;
}

// true when the layer ran on tensor cores; false = not covered here (caller falls back to the fp32 kernel)
static bool convlayer_tc_launch(ConvLayerTC* t, const ActView& x, const ConvDesc& c, long long n_items, float* y,
                                const float* resid, cudaStream_t st, cudaError_t* e) {
    *e = cudaSuccess;
    if (!t || x.is_u8 || x.sc != 1 || x.ch != c.cin || (x.sl % 4) || (x.sn % 4) ||
        ((reinterpret_cast<uintptr_t>(x.base) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(resid)) & 15))
        return false;                                  // 128-bit loads and the rows' bulk copies need 16-byte aligned bases
    auto it = t->layers.find(c.w);
    if (it == t->layers.end()) return false;
    if (n_items <= 0) return true;
    const cl::Geometry& g = it->second.g;
    cl::ConvTcArgs a;
    a.x = static_cast<const float*>(x.base); a.sn = x.sn; a.sl = x.sl;
    a.w = it->second.d_w; a.bias = c.b; a.y = y; a.resid = resid;
    a.lin = x.len; a.lout = c.out_len(x.len);
    a.cin = c.cin; a.cout = c.cout; a.ksz = c.k; a.stride = c.stride; a.relu = c.relu;
    a.nt = g.nt; a.stacked = g.stacked; a.kc = g.kc; a.n_chunks = c.cin / g.kc; a.n_arr = g.n_arr;
    a.sigma_min = g.sigma_min; a.max_delta = g.max_delta;
    for (int k = 0; k < cl::MAX_TAPS; ++k) { a.tap_arr[k] = g.tap_arr[k]; a.tap_delta[k] = g.tap_delta[k]; }
    a.pitch = g.pitch(a.lin, a.lout);
    a.trace = t->d_trace;
    const long long total = n_items * a.pitch;
    if (a.lout <= 0 || total + 128 + cl::ROWS_TAB >= 0x7fffffffLL) return false;      // 32-bit row arithmetic in the kernel
    a.n_items = (uint32_t)n_items; a.total_rows = (uint32_t)total;
    a.m_tiles = (uint32_t)((total + 127) / 128); a.n_tiles = (uint32_t)(c.cout / g.nt);
    const long long work = (long long)a.m_tiles * a.n_tiles;
    if (work >= 0x7fffffffLL) return false;
    a.acc_shift = g.acc_shift;
    a.res_ring = g.cps == 2 && resid && convlayer_tc_smem(t->mode, g) + cl::ring_bytes() <= cl::HALF_SM_SMEM;
    const unsigned grid = (unsigned)std::min<long long>(work, (long long)t->sm_count * g.cps);
    // one CTA per SM takes all of tensor memory: ask for more than half an SM's shared memory so that two never share one
    const size_t smem = g.cps == 2 ? convlayer_tc_smem(t->mode, g) + (a.res_ring ? cl::ring_bytes() : 0)
                                   : std::max(convlayer_tc_smem(t->mode, g), cl::MIN_SMEM);
    if (t->mode == 3) {
        if (g.cps == 2) cl::convlayer_tc_kernel<3, 2><<<grid, cl::THREADS, smem, st>>>(a);
        else cl::convlayer_tc_kernel<3, 1><<<grid, cl::THREADS, smem, st>>>(a);
    } else {
        if (g.cps == 2) cl::convlayer_tc_kernel<1, 2><<<grid, cl::THREADS, smem, st>>>(a);
        else cl::convlayer_tc_kernel<1, 1><<<grid, cl::THREADS, smem, st>>>(a);
    }
    *e = cudaGetLastError();
    return true;
}

static void convlayer_tc_destroy(ConvLayerTC* t) {
    if (!t) return;
    for (auto& kv : t->layers) if (kv.second.d_w) cudaFree(kv.second.d_w);
    delete t;
}

}  // namespace hello
