/*
 * hello_moe.h -- C ABI of the B200 (sm_100a) forward of HELLO's mixture-of-experts variant-calling DNN.
 *
 * The reference has no FFI for this path: the DNN is called as a Python object,
 *     network(featureDict, ref_segment)                      python/caller_calling.py:651-652
 *     dnn(tensors, numAllelesPerSite, numReadsPerAllele,
 *         reference_segments, numReadsPerSite)                python/MixtureOfExpertsDNNFast.py:128-134
 * and everything underneath is torch.nn on the CPU (python/MixtureOfExpertsAdvanced.py:161-252, 520-589).
 * This header is the boundary a binding for that call would use; hello_b200/_lib.py is the ctypes
 * binding, hello_b200/model.py mirrors the two Python signatures on top of it.
 *
 * Conventions
 *   - plain pointers and sizes only; every `const T* d_*` / `T* d_*` argument is DEVICE memory on the
 *     handle's device, every `h_*` argument is HOST memory;
 *   - the caller owns every buffer; forward() allocates nothing, never synchronises the stream and never
 *     calls back into the host; all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - functions return 0 on success or a negative hello_status; nothing throws across the ABI;
 *   - a handle is bound to one device and is not re-entrant (one forward at a time per handle).
 *
 * Ragged layout (CSR).  S sites, A alleles in total, R_t reads of technology t in total.
 *   site_allele_off[S+1]      alleles of site s are [off[s], off[s+1])        (numAllelesPerSite prefix sum)
 *   allele_read_off_t[A+1]    reads of allele a, technology t                 (numReadsPerAllele[t] prefix sum)
 *   every slot holds at least one row, as in the reference (an allele without support in a technology is
 *   one all-zero read: python/AlleleSearcherLite.py:245-247).
 */
#ifndef HELLO_MOE_H
#define HELLO_MOE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HELLO_MOE_ABI_VERSION 3

typedef enum hello_status {
    HELLO_OK = 0,
    HELLO_ERR_ARG = -1,        /* bad argument / inconsistent sizes */
    HELLO_ERR_BLOB = -2,       /* malformed weight blob */
    HELLO_ERR_CUDA = -3,       /* a CUDA runtime call failed (see hello_moe_last_error) */
    HELLO_ERR_WORKSPACE = -4,  /* workspace too small for even one site */
    HELLO_ERR_UNSUPPORTED = -5 /* topology the kernels do not cover */
} hello_status;

enum { HELLO_LAYOUT_RCL = 0,   /* reads as [R, C, L]  -- what MoEAttention.forward receives (.py:161-162)      */
       HELLO_LAYOUT_RLC = 1 }; /* reads as [R, L, C]  -- what the C++ encoder emits and featureDict carries
                                  (c++/src/AlleleSearcherLiteFiltered.cpp:1031-1180, caller_calling.py:633-639) */

enum { HELLO_META_NONE = 0,
       HELLO_META_SITE = 1,    /* architectures/meta_convolver.py: input = combined site features             */
       HELLO_META_REF = 2 };   /* architectures/meta_convolver_ref.py: input = one-hot reference segment      */

enum { HELLO_COMBINE_NONE = 0, /* no hybrid expert                                                            */
       HELLO_COMBINE_CONV = 1, /* combiner0 (allele level) and combiner1 (site level): MoEAttention, :193-219       */
       HELLO_COMBINE_SUM = 2 };/* legacy MoEMergedAdvanced with useAdditive and no ConvCombiners: allele features
                                  added, site frame = sum of that over the site's alleles (:408-436)              */

enum { HELLO_PREC_FP32 = 0,    /* fp32 FMA everywhere (CUDA cores)                                            */
       HELLO_PREC_BF16X3 = 1,  /* read convolver, compressor, combiner, xattn and meta_convolver on tcgen05:
                                  operands split into hi+lo bf16, three products, fp32 accumulate             */
       HELLO_PREC_BF16 = 2 };  /* same kernels with single bf16 operands, fp32 accumulate ("fast" mode)       */

/* Model wiring = which sub-networks MoEAttention holds (python/moe_attention_config_*.py). */
typedef struct hello_cfg {
    int32_t struct_size;       /* sizeof(hello_cfg), for forward compatibility */
    int32_t n_tech;            /* 1 (single technology) or 2 (hybrid) */
    int32_t read_channels[2];  /* 6, or 7 with the haplotag channel */
    int32_t xattn_present[3];  /* expert heads xattn0/1/2 */
    int32_t has_combiners;     /* HELLO_COMBINE_*: how the two technologies' features become the hybrid expert's input */
    int32_t meta_kind;         /* HELLO_META_* */
    int32_t feature_length;    /* 150 */
    int32_t precision;         /* HELLO_PREC_* */
    int32_t max_chunk_sites;   /* 0 = choose from the workspace size */
} hello_cfg;

typedef struct hello_moe hello_moe;

/* Inputs of one forward call. */
typedef struct hello_batch {
    int64_t n_sites;                     /* S */
    int64_t n_alleles;                   /* A */
    int64_t n_reads[2];                  /* R_0, R_1 (R_1 = 0 for single technology) */
    int32_t input_layout;                /* HELLO_LAYOUT_* of d_reads */
    int32_t reserved;
    const uint8_t* d_reads[2];           /* uint8 feature tensors [R_t, ., .] in `input_layout` */
    const int32_t* d_allele_read_off[2]; /* [A+1] each */
    const int32_t* h_allele_read_off[2]; /* host copies of the same arrays (used to plan chunks) */
    const int32_t* d_site_allele_off;    /* [S+1] */
    const int32_t* h_site_allele_off;    /* host copy */
    const float* d_ref_onehot;           /* [S, L, 5] fp32 one-hot reference segment (caller_calling.py:642-649);
                                            needed only for HELLO_META_REF, else may be NULL */
    const int32_t* d_allele_rank;        /* [A] tie-break rank of each allele inside its site (rank of the allele
                                            string in sorted order); NULL = use the allele's index in the site */
    const int64_t* d_pair_off;           /* [S+1] prefix sum of A_s*(A_s+1)/2 (genotype pairs per site) */
} hello_batch;

/* Outputs of one forward call (all device memory, caller-allocated). */
typedef struct hello_result {
    float* d_logits;      /* [3, A]  expert logits e0,e1,e2 (absent experts: 0) -- MoEAttention.forward :237-252 */
    float* d_meta;        /* [S, 3]  softmaxed expert weights; (1,0,0) when the model has no meta network        */
    float* d_pair_prob;   /* [4, P]  P = d_pair_off[S]; rows: mixed, P_e0, P_e1, P_e2 in the reference's pair
                                     order (i<=j, row-major) -- MoEMergedWrapperAdvanced.forward :559-584        */
    double* d_pair_mix64; /* [P]     float64 re-mix sum_e double(P_e)*double(meta_e) (prepareVcf.py:154-162);
                                     may be NULL */
    int32_t* d_best_pair; /* [S, 2]  argmax genotype (allele indices inside the site), ties broken as the
                                     reference does: greatest (allele_i, allele_j) key (caller_calling.py:702-705) */
    float* d_best_prob;   /* [S]     its probability */
    /* The final-call step that follows the network (prepareVcf.py:36-105,142-175; caller_calling.py:702-735).  All
     * three may be NULL.  Call k: 0 = the fp32 mixture above, 1..3 = expert k-1 alone, 4 = the float64 re-mix
     * ("mean" record); the "best" record of prepareVcf is call 1 + d_best_expert[s]. */
    int32_t* d_call_pair;   /* [S, 5, 2] argmax genotype of every call (same tie-break as d_best_pair)            */
    double* d_call_qual;    /* [S, 5]    QUAL = -10 log10(1 - min(p, 1 - 1e-8)), float64 (prepareVcf.py:60-62)     */
    int32_t* d_best_expert; /* [S]       np.argmax(meta) (prepareVcf.py:148,168)                                   */
} hello_result;

/* ABI version of the loaded library (== HELLO_MOE_ABI_VERSION). */
int hello_moe_abi_version(void);

/* Build a model from a packed weight blob (format: hello_b200/weights.py:pack_blob; weight-norm already folded,
 * replaces NNTools.WeightNormedConv1d/Linear recomputing g*v/|v| on every forward, python/NNTools.py:780-799).
 * Copies the weights to `device`. */
int hello_moe_create(const void* blob, size_t nbytes, const hello_cfg* cfg, int device, hello_moe** out);

void hello_moe_destroy(hello_moe* h);

/* Last error message of this handle (or of create() when h is NULL). Never NULL. */
const char* hello_moe_last_error(const hello_moe* h);

/* Workspace needed to run a chunk of the given size in one pass. forward() accepts any workspace that fits
 * at least one site and splits the batch into chunks of sites accordingly. */
size_t hello_moe_workspace_bytes(const hello_moe* h, int64_t n_reads0, int64_t n_reads1, int64_t n_alleles,
                                 int64_t n_sites);

/* The batched forward: MoEAttention.forward (python/MixtureOfExpertsAdvanced.py:161-252) followed by the
 * per-site genotype-pair enumeration, expert mixing and argmax of MoEMergedWrapperAdvanced.forward
 * (:527-589) and caller_calling.py:702-705. */
int hello_moe_forward(hello_moe* h, const hello_batch* in, const hello_result* out, void* d_workspace,
                      size_t workspace_bytes, void* stream);

/* The same forward for sites [site_begin, site_end) of a batch whose buffers describe ALL n_sites sites: every offset
 * array, read tensor and result buffer is the whole batch's, only the given sites are computed and only their result
 * slots written.  n_pairs = d_pair_off[n_sites] (row stride of d_pair_prob).  This is what a streaming caller uses:
 * upload the read rows of the next site range on a copy stream while this range computes (MoEEngine.forward_host). */
int hello_moe_forward_range(hello_moe* h, const hello_batch* in, const hello_result* out, int64_t site_begin,
                            int64_t site_end, int64_t n_pairs, void* d_workspace, size_t workspace_bytes, void* stream);

/* Stage timing (a measurement aid for bench.py): while enabled, forward() brackets the read-convolver stage of
 * every chunk (the dominant kernel) with CUDA events on the launching stream. collect() waits for those events and
 * returns the accumulated milliseconds and the number of bracketed regions since the previous collect. */
int hello_moe_profile_enable(hello_moe* h, int on);
int hello_moe_profile_collect(hello_moe* h, double* ms_read_conv, int64_t* n_regions);

/* Number of kernel launches enqueued by this handle since creation (bench.py reports the delta). */
int64_t hello_moe_launch_count(const hello_moe* h);

/* Test hook: run ONE sub-network on `n_items` items of length `lin`.
 * net_id: 0,1 read_convolver0/1 (input uint8 in `input_layout`); 2,3 compressor0/1; 4,5,6 xattn0/1/2 (input is the
 * already-combined 2a-s tensor); 7,8 combiner0/1 (input already concatenated); 9 meta.
 * Float inputs/outputs are channel-last [n, L, C] fp32. Returns output (channels, length) through out_c/out_l. */
int hello_moe_run_net(hello_moe* h, int net_id, const void* d_in, int64_t n_items, int32_t lin,
                      int32_t input_layout, float* d_out, int32_t* out_c, int32_t* out_l, void* d_workspace,
                      size_t workspace_bytes, void* stream);

/* Test hook for the fused tensor-core read convolver (HELLO_PREC_BF16X3 / HELLO_PREC_BF16 handles only): run it on
 * n_reads uint8 rows and, when d_dbg is not NULL, also dump the post-activation values of layer phase `phase`
 * (0-2 stem convs, 3-8 the six 32-channel convs, 9-10 the stride-2 block, 11-16 the six 64-channel convs) as
 * fp32 [ceil(n_reads/3), 512, 64]: row index = packed row of the 3-read group, 64 = channel slots.  Phases 0-1: pitch 160,
 * row = position, 16 channels.  Phases 2-8 (32-channel stage, run space-to-depth): pitch 40, row r holds positions 2r and 2r+1
 * as channel slots [0,32) and [32,64).  Phases 9-16: pitch 40, row = position, 64 channels. */
int hello_moe_readconv_debug(hello_moe* h, int tech, const uint8_t* d_reads, int64_t n_reads, int32_t input_layout,
                             int32_t phase, float* d_out, float* d_dbg, void* stream);

/* Test hook for the fused tensor-core head networks (HELLO_PREC_BF16X3 / HELLO_PREC_BF16 handles only; net_id 2,3
 * compressor, 4,5,6 xattn on an already combined 2a-s input, 9 meta_convolver): run the kernel on n_items fp32
 * channel-last items and, when d_dbg is not NULL, dump the post-activation values of layer phase `phase`
 * (0: 1x1 conv, 1-2: stride-2 block conv a / block output, 3-6: the two identity blocks) as fp32
 * [groups, 256, 256] (combiner, net_id 7,8, input [n,18,256]: phase 0 only = the 512-channel intermediate as
 * [groups of 6 items, 128, 512], pitch 20): a group is 6 (compressor) or 12 items, row = packed row of the group (pitch 40 / 20 in phase 0,
 * half of it afterwards), column = channel. d_out as for hello_moe_run_net. */
int hello_moe_headconv_debug(hello_moe* h, int net_id, const float* d_in, int64_t n_items, int32_t phase, float* d_out,
                             float* d_dbg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HELLO_MOE_H */
