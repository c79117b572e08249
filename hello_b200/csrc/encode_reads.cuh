// Read-feature encoder: AlleleSearcherLiteFiltered::computeFeaturesColoredSimple (reference:
// c++/src/AlleleSearcherLiteFiltered.cpp:1031-1180) for a whole batch of rows in one launch.
//
// One warp per output row.  The 150 window positions are spread over the lanes (position = lane + 32 k, 5 per lane)
// and kept in registers as one 64-bit word of up to 7 channel bytes each; the CIGAR is walked operation by operation
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
__device__ __forceinline__ unsigned long long set_byte(unsigned long long w, int track, uint32_t v) {
    const int sh = track * 8;
    return (w & ~(0xffull << sh)) | ((unsigned long long)(v & 0xffu) << sh);
}

__global__ void __launch_bounds__(WARPS * 32) encode_reads_kernel(const hello_encode_batch b, const Luts lut,
                                                                   uint8_t* __restrict__ out) {
    __shared__ __align__(16) uint8_t stage[WARPS][MAX_L * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * WARPS + warp;
    if (row >= b.n_rows) return;
    const int L = b.feature_length, C = b.channels;
    unsigned long long px[PER_LANE];
#pragma unroll
    for (int k = 0; k < PER_LANE; ++k) px[k] = 0ull;
    const int rid = b.d_row_read[row];
    if (rid >= 0) {
        const int site = b.d_row_site[row];
        const long long wstart = b.d_window_start[site], a0 = b.d_assembly_start[site], a1 = b.d_assembly_stop[site];
        const uint8_t* ref = b.d_reference + b.d_ref_off[site];
        const long long start = (a0 + a1) / 2 - (long long)(L / 2), end = start + L;
        const uint8_t* bases = b.d_bases + b.d_read_off[rid];
        const uint8_t* quals = b.d_quals + b.d_read_off[rid];
        const long long c0 = b.d_cigar_off[rid], c1 = b.d_cigar_off[rid + 1];
        long long rf = b.d_ref_start[rid], rd = 0;
        const uint32_t mq = lut.map_q[b.d_mapq[rid]];
        const uint32_t sc = b.d_orientation[rid] > 0 ? 70u : 240u;
        const uint32_t hp_raw = C == 7 ? b.d_hp[rid] : 0u;
        const uint32_t hc = hp_raw == 1 ? 120u : (hp_raw == 2 ? 240u : 0u);
        // tracks every coloured position gets: mapq, strand, (hp)
        unsigned long long common = ((unsigned long long)mq << (8 * T_READ_MAPQ)) | ((unsigned long long)sc << (8 * T_ORIENT));
        if (C == 7) common |= (unsigned long long)hc << (8 * T_HP);
        auto pos_color = [&](long long pos) -> uint32_t {               // :1007-1015 (unsigned comparison in the C++)
            const unsigned long long p = (unsigned long long)(pos - wstart);
            return ((unsigned long long)(a0 - wstart) <= p && p < (unsigned long long)(a1 - wstart)) ? 240u : 70u;
        };
        for (long long ci = c0; ci < c1; ++ci) {
            const uint32_t cg = __ldg(b.d_cigars + ci);
            const uint32_t op = cg & 15u;
            const long long len = cg >> 4;
            if (op == 0 || op == 7 || op == 8) {                         // M, =, X
                if (rf < end && rf + len > start) {
#pragma unroll
                    for (int k = 0; k < PER_LANE; ++k) {
                        const int f = lane + 32 * k;
                        const long long pos = start + f;
                        if (f < L && pos >= rf && pos < rf + len) {
                            const long long j = pos - rf;
                            unsigned long long w = common;
                            w |= (unsigned long long)base_color(__ldg(bases + rd + j)) << (8 * T_READ_BASE);
                            w |= (unsigned long long)base_color(__ldg(ref + (pos - wstart))) << (8 * T_REF_BASE);
                            w |= (unsigned long long)lut.base_q[__ldg(quals + rd + j)] << (8 * T_READ_QUAL);
                            w |= (unsigned long long)pos_color(pos) << (8 * T_POSITION);
                            px[k] = w;
                        }
                    }
                }
                rf += len; rd += len;
            } else if (op == 2 || op == 3) {                             // D (falls through into N)
                if (op == 2 && start <= rf - 1 && rf - 1 < end) {
                    const uint32_t q0 = rd > 0 ? lut.base_q[__ldg(quals + rd - 1)] : 0u;
#pragma unroll
                    for (int k = 0; k < PER_LANE; ++k) {
                        const int f = lane + 32 * k;
                        const long long pos = start + f;
                        if (f < L && pos >= rf - 1 && pos < rf + len) {
                            // reference base, mapq, strand, position marker, hp; read base / quality keep their value
                            unsigned long long w = px[k];
                            w = set_byte(w, T_REF_BASE, base_color(__ldg(ref + (pos - wstart))));
                            w = set_byte(w, T_READ_MAPQ, mq);
                            w = set_byte(w, T_ORIENT, sc);
                            w = set_byte(w, T_POSITION, pos_color(pos));
                            if (C == 7) w = set_byte(w, T_HP, hc);
                            if (pos == rf - 1) {                          // '*' and the quality of the base before the gap
                                w = set_byte(w, T_READ_BASE, 0u);
                                w = set_byte(w, T_READ_QUAL, q0);
                            }
                            px[k] = w;
                        }
                    }
                }
                rf += len;
            } else if (op == 1 || op == 4) {                             // I (falls through into S)
                if (op == 1 && start <= rf - 1 && rf - 1 < end) {
                    const long long lo = rd > 0 ? rd - 1 : rd, hi = rd + len;
                    uint32_t qmin = 255u;                                // min over the base before and the inserted bases
                    for (long long t = lo + lane; t < hi; t += 32) qmin = min(qmin, (uint32_t)__ldg(quals + t));
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) qmin = min(qmin, __shfl_xor_sync(0xffffffffu, qmin, d));
                    const long long pos = rf - 1;
                    const int f = (int)(pos - start);
                    unsigned long long w = common;
                    w |= (unsigned long long)base_color(__ldg(ref + (pos - wstart))) << (8 * T_REF_BASE);
                    w |= (unsigned long long)lut.base_q[qmin] << (8 * T_READ_QUAL);
                    w |= (unsigned long long)pos_color(pos) << (8 * T_POSITION);
#pragma unroll
                    for (int k = 0; k < PER_LANE; ++k) px[k] = (f == lane + 32 * k) ? w : px[k];   // selects keep px in registers
                }
                rd += len;
            }
            // H, P, B: no case in the reference's switch
        }
    }
    // stage the row [L][C] and write it out with coalesced 32-bit words
    uint8_t* st = stage[warp];
#pragma unroll
    for (int k = 0; k < PER_LANE; ++k) {
        const int f = lane + 32 * k;
        if (f < L) {
            for (int c = 0; c < C; ++c) st[f * C + c] = (uint8_t)(px[k] >> (8 * c));
        }
    }
    __syncwarp();
    const int row_bytes = L * C;
    uint8_t* dst = out + row * row_bytes;
    if (((row_bytes | (int)((uintptr_t)dst & 3)) & 3) == 0) {
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(st);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
        for (int i = lane; i < row_bytes / 4; i += 32) d4[i] = s4[i];
    } else {
        for (int i = lane; i < row_bytes; i += 32) dst[i] = st[i];
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
