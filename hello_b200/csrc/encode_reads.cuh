// Read-feature encoder: AlleleSearcherLiteFiltered::computeFeaturesColoredSimple (reference:
// c++/src/AlleleSearcherLiteFiltered.cpp:1031-1180) for a whole batch of rows in one launch.
//
// One warp per output row.  The 150 window positions are spread over the lanes (position = lane + 32 k, 5 per lane)
// and kept in registers as two 32-bit words of channel bytes each; the CIGAR is walked operation by operation
// by the whole warp (the operations are warp-uniform), every lane colouring the window positions it owns, so later
// operations overwrite earlier ones exactly as the sequential C++ does.  The finished row is staged in shared memory
// and written with coalesced 32-bit stores (a row is 900 or 1050 contiguous bytes).  HBM-bound byte work: ~0.4 KB of
// read data in, 0.9 KB out per row.
#pragma once
#include <string>

#include "../../include/hello_encode.h"
#include "../../include/hello_moe.h"
#include "common.cuh"

namespace hello {
namespace enc {

constexpr int WARPS = 8, MAX_L = 160, MAX_C = 7, PER_LANE = MAX_L / 32;
enum { T_READ_BASE = 0, T_REF_BASE, T_READ_QUAL, T_READ_MAPQ, T_ORIENT, T_POSITION, T_HP };

struct Luts { uint8_t base_q[256]; uint8_t map_q[256]; };

__device__ __forceinline__ uint32_t base_color(uint8_t b) {            // :971-984
    return b == 'A' ? 250u : b == 'G' ? 180u : b == 'T' ? 100u : b == 'C' ? 30u : 0u;
}

// Per window position f the tracks split into what depends only on (site, f) -- reference base and the allele-span marker,
// computed when the warp moves to a new site --, what depends on the read -- mapq, strand, hp, one value per row -- and what
// the CIGAR walk decides: read base and base quality.  Low word = tracks 0-3 (read base, ref base, quality, mapq); the
// high word (tracks 4-6) of a position is either zero (never touched) or the row's constant, so it is kept as one bit.
// All window arithmetic is 32-bit and relative to the window start.  A warp encodes ROWS_PER_WARP consecutive rows: rows
// come in site order, so the per-site part (three 64-bit records, the clamps, five reference bases) is paid once per
// site and warp instead of once per row.
constexpr int ROWS_PER_WARP = 16;

__global__ void __launch_bounds__(WARPS * 32, 8) encode_reads_kernel(const hello_encode_batch b, const Luts lut,
                                                                   uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t stage[WARPS][MAX_L * 8];
    __shared__ uint8_t s_base[256], s_qual[256];
    s_base[threadIdx.x] = (uint8_t)base_color((uint8_t)threadIdx.x);
    s_qual[threadIdx.x] = lut.base_q[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = b.feature_length, C = b.channels;
    const int row_bytes = L * C;
    long long row = ((long long)blockIdx.x * WARPS + warp) * ROWS_PER_WARP;
    const long long row_end = min(row + ROWS_PER_WARP, (long long)b.n_rows);
    // per-site state
    int cur_site = -1;
    uint32_t site_lo[PER_LANE];                 // reference-base colour << 8 of the lane's positions
    uint32_t span = 0u;                         // bit k: position lane + 32 k lies inside the allele span
    long long start = 0;
    for (; row < row_end; ++row) {
        uint32_t lo[PER_LANE];
        uint32_t touched = 0u, hi_row = 0u;
#pragma unroll
        for (int k = 0; k < PER_LANE; ++k) lo[k] = 0u;
        const int rid = b.d_row_read[row];
        if (rid >= 0) {
            const int site = b.d_row_site[row];
            if (site != cur_site) {
                cur_site = site;
                const long long wstart = b.d_window_start[site], a0 = b.d_assembly_start[site], a1 = b.d_assembly_stop[site];
                start = (a0 + a1) / 2 - (long long)(L / 2);
                const long long ref_off = b.d_ref_off[site], ref_len = b.d_ref_off[site + 1] - ref_off;
                const uint8_t* refw = b.d_reference + ref_off + (start - wstart);           // reference base of window position f
                // everything below is window-relative and clamped to +-2^30, so the per-position tests are 32-bit
                auto clamp30 = [](long long v) { return (int)max(-(1ll << 30), min(1ll << 30, v)); };
                const int ref_lo = clamp30(wstart - start), ref_hi = clamp30(wstart - start + ref_len);   // valid f range of refw
                // PositionColor (:1007-1015) compares unsigned offsets from windowStart
                const int p_lo = a0 >= wstart ? clamp30(a0 - start) : (1 << 30), p_hi = a1 >= wstart ? clamp30(a1 - start) : (1 << 30);
                span = 0u;
#pragma unroll
                for (int k = 0; k < PER_LANE; ++k) {
                    const int f = lane + 32 * k;
                    site_lo[k] = (f >= ref_lo && f < ref_hi) ? ((uint32_t)s_base[__ldg(refw + f)] << 8) : 0u;
                    span |= (f >= p_lo && f < p_hi) ? (1u << k) : 0u;
                }
            }
            auto clamp30 = [](long long v) { return (int)max(-(1ll << 30), min(1ll << 30, v)); };
            const long long r_off = b.d_read_off[rid];
            const uint8_t* bases = b.d_bases + r_off;
            const uint8_t* quals = b.d_quals + r_off;
            const long long c0 = b.d_cigar_off[rid];
            const int n_ops = (int)(b.d_cigar_off[rid + 1] - c0);
            const uint32_t* cigar = b.d_cigars + c0;
            const uint32_t hp_raw = C == 7 ? b.d_hp[rid] : 0u;
            hi_row = (b.d_orientation[rid] > 0 ? 70u : 240u) | ((hp_raw == 1 ? 120u : (hp_raw == 2 ? 240u : 0u)) << 16);
            const uint32_t mq = (uint32_t)lut.map_q[b.d_mapq[rid]] << 24;
            // window-relative reference cursor, clamped far outside the window instead of overflowing
            int rf = clamp30(b.d_ref_start[rid] - start);
            int rd = 0;
            for (int ci = 0; ci < n_ops; ++ci) {
                const uint32_t cg = __ldg(cigar + ci);
                const uint32_t op = cg & 15u;
                const int len = (int)(cg >> 4);
                if (op == 0 || op == 7 || op == 8) {                         // M, =, X
                    if (rf < L && rf + len > 0) {
                        const uint8_t* bp = bases + (rd - rf) + lane;        // position f reads base rd + (f - rf)
                        const uint8_t* qp = quals + (rd - rf) + lane;
#pragma unroll
                        for (int k = 0; k < PER_LANE; ++k) {
                            const int f = lane + 32 * k;
                            if (f < L && (unsigned)(f - rf) < (unsigned)len) {
                                lo[k] = site_lo[k] | mq | s_base[__ldg(bp + 32 * k)] | ((uint32_t)s_qual[__ldg(qp + 32 * k)] << 16);
                                touched |= 1u << k;
                            }
                        }
                    }
                    rf = min(rf + len, 1 << 30); rd += len;
                } else if (op == 2 || op == 3) {                             // D (falls through into N)
                    if (op == 2 && rf - 1 >= 0 && rf - 1 < L) {
                        const uint32_t q0 = rd > 0 ? s_qual[__ldg(quals + rd - 1)] : 0u;
#pragma unroll
                        for (int k = 0; k < PER_LANE; ++k) {
                            const int f = lane + 32 * k;
                            if (f < L && f >= rf - 1 && f < rf + len) {
                                // reference base, mapq, strand, position marker, hp; read base / quality keep their value,
                                // except at the base before the gap: '*' and that base's quality
                                lo[k] = f == rf - 1 ? (site_lo[k] | mq | (q0 << 16)) : ((lo[k] & 0x00ff00ffu) | site_lo[k] | mq);
                                touched |= 1u << k;
                            }
                        }
                    }
                    rf = min(rf + len, 1 << 30);
                } else if (op == 1 || op == 4) {                             // I (falls through into S)
                    if (op == 1 && rf - 1 >= 0 && rf - 1 < L) {
                        const int q_lo = rd > 0 ? rd - 1 : rd, q_hi = rd + len;
                        uint32_t qmin = 255u;                                // min over the base before and the inserted bases
                        for (int t = q_lo + lane; t < q_hi; t += 32) qmin = min(qmin, (uint32_t)__ldg(quals + t));
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) qmin = min(qmin, __shfl_xor_sync(0xffffffffu, qmin, d));
                        const uint32_t qc = (uint32_t)s_qual[qmin] << 16;
#pragma unroll
                        for (int k = 0; k < PER_LANE; ++k) {                 // selects keep the arrays in registers
                            const bool hit = rf - 1 == lane + 32 * k;
                            lo[k] = hit ? (site_lo[k] | mq | qc) : lo[k];
                            touched |= hit ? (1u << k) : 0u;
                        }
                    }
                    rd += len;
                }
                // H, P, B: no case in the reference's switch
            }
        }
        // tracks 4-6 of a touched position: strand, allele-span marker, hp
        uint8_t* dst = out + row * row_bytes;
        if (C == 6 && (((uintptr_t)dst) & 1) == 0) {
            // 6-byte records of consecutive lanes are contiguous: three 16-bit stores per position, each warp instruction
            // covering a third of a 192-byte span; L2 merges them, nothing is staged.
            uint16_t* d2 = reinterpret_cast<uint16_t*>(dst) + lane * 3;
#pragma unroll
            for (int k = 0; k < PER_LANE; ++k) {
                if (lane + 32 * k < L) {
                    const uint32_t hi = (touched >> k) & 1u ? (hi_row | (((span >> k) & 1u ? 240u : 70u) << 8)) : 0u;
                    d2[96 * k] = (uint16_t)lo[k];
                    d2[96 * k + 1] = (uint16_t)(lo[k] >> 16);
                    d2[96 * k + 2] = (uint16_t)hi;
                }
            }
            continue;
        }
        // otherwise stage the row [L][C] in shared memory and write it out with coalesced 32-bit words
        uint8_t* st = stage[warp];
        __syncwarp();                                                        // the previous row has left the staging buffer
#pragma unroll
        for (int k = 0; k < PER_LANE; ++k) {
            const int f = lane + 32 * k;
            if (f < L) {
                const uint32_t hi = (touched >> k) & 1u ? (hi_row | (((span >> k) & 1u ? 240u : 70u) << 8)) : 0u;
                uint8_t* q = st + f * C;
                q[0] = (uint8_t)lo[k]; q[1] = (uint8_t)(lo[k] >> 8); q[2] = (uint8_t)(lo[k] >> 16); q[3] = (uint8_t)(lo[k] >> 24);
                q[4] = (uint8_t)hi; q[5] = (uint8_t)(hi >> 8);
                if (C == 7) q[6] = (uint8_t)(hi >> 16);
            }
        }
        __syncwarp();
        if (((row_bytes | (int)((uintptr_t)dst & 3)) & 3) == 0) {
            const uint32_t* s4 = reinterpret_cast<const uint32_t*>(st);
            uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
            for (int i = lane; i < row_bytes / 4; i += 32) d4[i] = s4[i];
        } else {
            for (int i = lane; i < row_bytes; i += 32) dst[i] = st[i];
        }
    }
}

inline Luts make_luts() {
    Luts l;
    for (int q = 0; q < 256; ++q) {
        // BaseQualityColor / MappingQualityColor (:987-998): float capped, double arithmetic, truncation
        float cb = (float)std::min(q, 40), cm = (float)std::min(q, 60);
        l.base_q[q] = (uint8_t)int(254 * (1.0 * cb / 40));
        l.map_q[q] = (uint8_t)int(254 * (1.0 * cm / 60));
    }
    return l;
}

}  // namespace enc
}  // namespace hello
