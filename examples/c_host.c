/*
 * A host program in plain C that binds the C ABI of include/hello_moe.h directly -- no Python, no torch: what a
 * non-Python caller of HELLO's DNN (the drop-in boundary) links against.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_host.c -o examples/c_host \
 *       -L hello_b200 -lhello_moe -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/hello_b200
 *   examples/c_host case.bin out.bin
 *
 * case.bin (little endian; written by tools/export_case.py from any model + pileups):
 *   char magic[8] = "HELLOCAS"; hello_cfg cfg; int64 blob_bytes, S, A, R0, R1;
 *   blob[blob_bytes]; reads0 u8 [R0,150,C0]; reads1 u8 [R1,150,C1];
 *   allele_read_off0 i32 [A+1]; allele_read_off1 i32 [A+1] (only when R1 > 0); site_allele_off i32 [S+1];
 *   ref_onehot f32 [S,150,5]
 * out.bin: int64 S, A, P; logits f32 [3,A]; meta f32 [S,3]; pair_prob f32 [4,P]; best_pair i32 [S,2]; best_prob f32 [S];
 *          call_qual f64 [S,5]
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hello_moe.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static void* slurp(FILE* f, size_t n) {
    void* p = malloc(n ? n : 1);
    if (!p || fread(p, 1, n, f) != n) { fprintf(stderr, "short read (%zu bytes)\n", n); exit(3); }
    return p;
}

static void* to_device(const void* h, size_t n) {
    void* d = NULL;
    if (cudaMalloc(&d, n ? n : 16) != cudaSuccess || cudaMemcpy(d, h, n, cudaMemcpyHostToDevice) != cudaSuccess) {
        fprintf(stderr, "upload of %zu bytes failed\n", n);
        exit(4);
    }
    return d;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s case.bin out.bin\n", argv[0]); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    char magic[8];
    hello_cfg cfg;
    int64_t hdr[5];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "HELLOCAS", 8) != 0 || fread(&cfg, sizeof cfg, 1, f) != 1 ||
        fread(hdr, sizeof hdr, 1, f) != 1) { fprintf(stderr, "bad case file\n"); return 1; }
    const int64_t blob_bytes = hdr[0], S = hdr[1], A = hdr[2], R[2] = {hdr[3], hdr[4]};
    const int L = cfg.feature_length;
    if (hello_moe_abi_version() != HELLO_MOE_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }

    void* blob = slurp(f, (size_t)blob_bytes);
    uint8_t* reads[2] = {NULL, NULL};
    int32_t* aro[2] = {NULL, NULL};
    for (int t = 0; t < 2; ++t) reads[t] = (uint8_t*)slurp(f, (size_t)R[t] * L * (t < cfg.n_tech ? cfg.read_channels[t] : 0));
    for (int t = 0; t < cfg.n_tech; ++t) aro[t] = (int32_t*)slurp(f, (size_t)(A + 1) * 4);
    int32_t* sao = (int32_t*)slurp(f, (size_t)(S + 1) * 4);
    float* onehot = (float*)slurp(f, (size_t)S * L * 5 * 4);
    fclose(f);

    int64_t* pair_off = (int64_t*)malloc((size_t)(S + 1) * 8);
    pair_off[0] = 0;
    for (int64_t s = 0; s < S; ++s) {
        const int64_t n = sao[s + 1] - sao[s];
        pair_off[s + 1] = pair_off[s] + n * (n + 1) / 2;           /* unordered genotype pairs, i <= j */
    }
    const int64_t P = pair_off[S];

    cfg.struct_size = (int32_t)sizeof cfg;
    hello_moe* h = NULL;
    int rc = hello_moe_create(blob, (size_t)blob_bytes, &cfg, 0, &h);
    if (rc != HELLO_OK) { fprintf(stderr, "hello_moe_create: %d %s\n", rc, hello_moe_last_error(NULL)); return 5; }

    hello_batch in;
    memset(&in, 0, sizeof in);
    in.n_sites = S; in.n_alleles = A; in.input_layout = HELLO_LAYOUT_RLC;
    for (int t = 0; t < cfg.n_tech; ++t) {
        in.n_reads[t] = R[t];
        in.d_reads[t] = (const uint8_t*)to_device(reads[t], (size_t)R[t] * L * cfg.read_channels[t]);
        in.d_allele_read_off[t] = (const int32_t*)to_device(aro[t], (size_t)(A + 1) * 4);
        in.h_allele_read_off[t] = aro[t];
    }
    in.d_site_allele_off = (const int32_t*)to_device(sao, (size_t)(S + 1) * 4);
    in.h_site_allele_off = sao;
    in.d_ref_onehot = (const float*)to_device(onehot, (size_t)S * L * 5 * 4);
    in.d_pair_off = (const int64_t*)to_device(pair_off, (size_t)(S + 1) * 8);

    hello_result out;
    memset(&out, 0, sizeof out);
    CK(cudaMalloc((void**)&out.d_logits, (size_t)(3 * A + 1) * 4));
    CK(cudaMalloc((void**)&out.d_meta, (size_t)(3 * S + 1) * 4));
    CK(cudaMalloc((void**)&out.d_pair_prob, (size_t)(4 * P + 1) * 4));
    CK(cudaMalloc((void**)&out.d_pair_mix64, (size_t)(P + 1) * 8));
    CK(cudaMalloc((void**)&out.d_best_pair, (size_t)(2 * S + 1) * 4));
    CK(cudaMalloc((void**)&out.d_best_prob, (size_t)(S + 1) * 4));
    CK(cudaMalloc((void**)&out.d_call_pair, (size_t)(10 * S + 1) * 4));
    CK(cudaMalloc((void**)&out.d_call_qual, (size_t)(5 * S + 1) * 8));
    CK(cudaMalloc((void**)&out.d_best_expert, (size_t)(S + 1) * 4));

    size_t ws_bytes = hello_moe_workspace_bytes(h, R[0], R[1], A, S);
    if (ws_bytes > ((size_t)2 << 30)) ws_bytes = (size_t)2 << 30;   /* a smaller workspace only means more chunks */
    void* ws = NULL;
    CK(cudaMalloc(&ws, ws_bytes));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    rc = hello_moe_forward(h, &in, &out, ws, ws_bytes, (void*)st);
    if (rc != HELLO_OK) { fprintf(stderr, "hello_moe_forward: %d %s\n", rc, hello_moe_last_error(h)); return 6; }
    CK(cudaStreamSynchronize(st));

    FILE* o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 1; }
    int64_t oh[3] = {S, A, P};
    fwrite(oh, sizeof oh, 1, o);
#define DUMP(ptr, count, type) do { size_t n_ = (size_t)(count) * sizeof(type); void* h_ = malloc(n_ ? n_ : 1); \
        CK(cudaMemcpy(h_, ptr, n_, cudaMemcpyDeviceToHost)); fwrite(h_, 1, n_, o); free(h_); } while (0)
    DUMP(out.d_logits, 3 * A, float);
    DUMP(out.d_meta, 3 * S, float);
    DUMP(out.d_pair_prob, 4 * P, float);
    DUMP(out.d_best_pair, 2 * S, int32_t);
    DUMP(out.d_best_prob, S, float);
    DUMP(out.d_call_qual, 5 * S, double);
    fclose(o);
    printf("c_host: %lld sites, %lld alleles, %lld genotype pairs scored; %lld kernel launches\n", (long long)S, (long long)A,
           (long long)P, (long long)hello_moe_launch_count(h));
    hello_moe_destroy(h);
    return 0;
}
