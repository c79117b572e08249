/*
 * hello_encode.h -- C ABI of the B200 (sm_100a) read-feature encoder: the step immediately before the network.
 *
 * Replaces AlleleSearcherLiteFiltered::computeFeaturesColoredSimple (c++/src/AlleleSearcherLiteFiltered.cpp:1031-1180;
 * colours :971-1027, constants :360-384), which the reference reaches through Boost.Python from
 * python/trainDataTools.py / caller_calling.py:612-640 once per (site, allele, technology) and whose uint8
 * [reads, 150, channels] arrays become the network's featureDict.  Here ONE launch encodes the rows of a whole batch
 * of sites straight into the [R, L, C] tensor hello_moe_forward takes (HELLO_LAYOUT_RLC), so the pileups never
 * cross PCIe as 900-byte rows: the host ships the aligned reads (bases, qualities, CIGARs), ~2.5x fewer bytes.
 *
 * Same conventions as hello_moe.h: plain pointers and sizes, `d_*` = device memory, the caller owns every buffer,
 * nothing is allocated or synchronised, work is enqueued on `stream`, 0 or a negative hello_status is returned.
 *
 * Row r of the output is read d_row_read[r] of site d_row_site[r]; d_row_read[r] < 0 asks for the all-zero row the
 * reference returns for an allele without support in a technology (:1037-1043).  The caller lists the rows in the
 * order the network wants them (site -> allele -> supporting reads, supports[allele] iteration order, reads of the
 * other technology skipped: numReadsSupportingAlleleStrict, :958-968).
 */
#ifndef HELLO_ENCODE_H
#define HELLO_ENCODE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hello_encode_batch {
    int64_t n_rows;                  /* rows to produce */
    int32_t feature_length;          /* 150 (python/call.py:187) */
    int32_t channels;                /* 6, or 7 = include_hp_tags */
    const int32_t* d_row_read;       /* [n_rows] read index, or -1 = all-zero row */
    const int32_t* d_row_site;       /* [n_rows] site index */
    /* reads (all sites back to back) */
    const int64_t* d_read_off;       /* [n_reads+1] offsets into d_bases / d_quals */
    const uint8_t* d_bases;          /* ASCII read bases */
    const uint8_t* d_quals;          /* phred qualities (0..255) */
    const int64_t* d_cigar_off;      /* [n_reads+1] offsets into d_cigars */
    const uint32_t* d_cigars;        /* BAM encoding: length << 4 | op  (op: 0 M, 1 I, 2 D, 3 N, 4 S, 5 H, 6 P, 7 =, 8 X) */
    const int64_t* d_ref_start;      /* [n_reads] reference position of the first aligned base */
    const uint8_t* d_mapq;           /* [n_reads] */
    const int8_t* d_orientation;     /* [n_reads] > 0 forward strand */
    const uint8_t* d_hp;             /* [n_reads] haplotag 0/1/2 (read only when channels == 7; may be NULL otherwise) */
    /* sites */
    const int64_t* d_ref_off;        /* [n_sites+1] offsets into d_reference */
    const uint8_t* d_reference;      /* ASCII reference windows; window of site s starts at position d_window_start[s] */
    const int64_t* d_window_start;   /* [n_sites] */
    const int64_t* d_assembly_start; /* [n_sites] allele span [assemblyStart, assemblyStop): the feature window is */
    const int64_t* d_assembly_stop;  /* [n_sites] centred on (assemblyStart + assemblyStop) / 2 (:1047-1050)        */
} hello_encode_batch;

/* d_out: uint8 [n_rows, feature_length, channels] (HELLO_LAYOUT_RLC).  Every byte of d_out is written. */
int hello_encode_reads(const hello_encode_batch* batch, uint8_t* d_out, void* stream);

/* Last error message of hello_encode_reads on this thread. Never NULL. */
const char* hello_encode_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* HELLO_ENCODE_H */
