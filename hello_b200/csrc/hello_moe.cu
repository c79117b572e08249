// hello_moe.cu -- C ABI (include/hello_moe.h) and host orchestration of the B200 MoE forward.
//
// Mirrors MoEAttention.forward (reference: python/MixtureOfExpertsAdvanced.py:161-252) followed by the per-site
// tail of MoEMergedWrapperAdvanced.forward (:527-589): read convolver per technology -> segmented sum over the
// reads of each allele -> compressor -> segmented sum over the alleles of each site -> 2a-s expert heads /
// combiners / meta gate -> sigmoid, genotype-pair probabilities, expert mixing, argmax.
// The batch is processed in chunks of whole sites sized to the caller's workspace; nothing is allocated here.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/hello_moe.h"
#include "common.cuh"
#include "conv_fp32.cuh"
#include "misc_kernels.cuh"
#include "readconv_tc.cuh"
#include "headconv_tc.cuh"
#include "combconv_tc.cuh"
#include "convlayer_tc.cuh"
#include "encode_reads.cuh"

using namespace hello;

namespace {

constexpr int N_NETS = 10;
enum NetId { NET_RC0 = 0, NET_RC1, NET_CMP0, NET_CMP1, NET_X0, NET_X1, NET_X2, NET_CB0, NET_CB1, NET_META };

thread_local std::string g_create_error;

struct Arena {
    char* base = nullptr;
    size_t cap = 0, top = 0, high = 0;
    bool dry = false, overflow = false;
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~size_t(255);
        char* p = base + top;
        top += bytes;
        high = std::max(high, top);
        if (!dry && top > cap) overflow = true;
        return p;
    }
    float* allocf(long long n) { return static_cast<float*>(alloc(size_t(n) * sizeof(float))); }
    size_t mark() const { return top; }
    void release(size_t m) { top = m; }
};

}  // namespace

constexpr size_t MAX_PROFILE_REGIONS = 4096;   // regions bracketed between two hello_moe_profile_collect calls
constexpr long long SMALL_BATCH_READS = 4 * 148 * 9;   // up to four work items per SM: see the read-convolver stage

struct hello_moe {
    hello_cfg cfg;
    int device = 0;
    float* d_weights = nullptr;
    std::vector<float> h_weights;    // host copy of the folded fp32 weights (packed to bf16 by the tensor-core path)
    size_t n_floats = 0;
    std::vector<LayerDesc> nets[N_NETS];
    std::string err;
    int64_t launches = 0;
    int read_len = 0, read_ch = 0;   // read convolver output (36, 64)
    int comp_len = 0, comp_ch = 0;   // compressor output (18, 128)
    ReadConvTC* tc[2] = {nullptr, nullptr};
    HeadConvTC* head[N_NETS] = {};   // fused tcgen05 compressor / xattn / meta_convolver (tensor-core precisions)
    CombConvTC* comb[2] = {nullptr, nullptr};   // fused tcgen05 combiner0 / combiner1
    ConvLayerTC* layer_tc = nullptr;            // generic tcgen05 layers for what the fused kernels do not cover
    // Sub-networks with an addendum (two more residual blocks stacked by the reference's build_on_top,
    // MixtureOfExpertsAdvancedXferLearning.py:94-183): the fused kernel runs the original layers, `tail` the added ones.
    std::vector<LayerDesc> tail[N_NETS];
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;            // pairs (start, stop), created lazily, at most MAX_PROFILE_REGIONS pairs
    bool small_call[2] = {false, false};         // this call scores at most SMALL_BATCH_READS reads of the technology
    size_t ev_used = 0;
};

namespace {

bool net_out_shape(const std::vector<LayerDesc>& net, int lin, int cin, int* lout, int* cout) {
    int l = lin, c = cin;
    for (const LayerDesc& L : net) {
        switch (L.kind) {
            case KIND_CONV: if (L.a.cin != c) return false; l = L.a.out_len(l); c = L.a.cout; break;
            case KIND_MAXPOOL: l = (l - L.a.k) / L.a.stride + 1; break;
            case KIND_RES:
                if (L.a.cin != c || L.b.cin != L.a.cout) return false;
                l = L.a.out_len(l); c = L.b.cout; break;
            case KIND_GAP_LINEAR: if (L.a.cin != c) return false; l = 1; c = L.a.cout; break;
            default: return false;
        }
        if (l <= 0) return false;
    }
    *lout = l; *cout = c;
    return true;
}

long long net_max_elems(const std::vector<LayerDesc>& net, int lin) {
    long long mx = 0;
    int l = lin;
    for (const LayerDesc& L : net) {
        if (L.kind == KIND_CONV) { l = L.a.out_len(l); mx = std::max(mx, (long long)l * L.a.cout); }
        else if (L.kind == KIND_MAXPOOL) { l = (l - L.a.k) / L.a.stride + 1; }
        else if (L.kind == KIND_RES) { l = L.a.out_len(l); mx = std::max(mx, (long long)l * L.b.cout); }
    }
    return mx;
}

struct Runner {
    hello_moe* h;
    Arena* arena;
    cudaStream_t st;
    bool dry;
    long long pair_total = 0;   // P = total genotype pairs of the batch (row stride of d_pair_prob)
    int status = HELLO_OK;

    bool fail(int code, const std::string& msg) {
        if (status == HELLO_OK) { status = code; h->err = msg; }
        return false;
    }
    bool check(cudaError_t e, const char* what) {
        if (e != cudaSuccess) return fail(HELLO_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
        return true;
    }
    bool launched(const char* what) {
        h->launches++;
        return check(cudaGetLastError(), what);
    }
    static unsigned grid_for(long long total, int block) {
        long long g = (total + block - 1) / block;
        return (unsigned)std::min<long long>(std::max<long long>(g, 1), 148LL * 64);
    }

    bool conv(const ActView& x, const ConvDesc& c, long long n, float* y, const float* resid) {
        if (dry || n == 0) return true;
        h->launches++;
        cudaError_t e;
        if (convlayer_tc_launch(h->layer_tc, x, c, n, y, resid, st, &e)) return check(e, "convlayer_tc");
        return check(launch_conv(x, c, n, y, resid, st), "conv1d_fp32");
    }

    struct GapOut { float* out; long long stride; int softmax; };

    // Run one sub-network on n items. Intermediate activations rotate through four slabs taken from the arena;
    // the last layer writes to `out` (feature nets) or through `gap` (nets ending in the pooled linear head).
    bool run_net(const std::vector<LayerDesc>& net, ActView cur, long long n, float* out, const GapOut* gap) {
        if (net.empty()) return fail(HELLO_ERR_UNSUPPORTED, "sub-network missing from the weight blob");
        const size_t m = arena->mark();
        const long long slab_elems = net_max_elems(net, cur.len) * n;
        float* slab[4];
        bool used[4] = {false, false, false, false};
        for (int i = 0; i < 4; ++i) slab[i] = arena->allocf(slab_elems);
        if (arena->overflow) return fail(HELLO_ERR_WORKSPACE, "workspace overflow");
        auto grab = [&]() { for (int i = 0; i < 4; ++i) if (!used[i]) { used[i] = true; return i; } return -1; };
        int cur_slab = -1;
        const size_t n_layers = net.size();
        for (size_t li = 0; li < n_layers && status == HELLO_OK; ++li) {
            const LayerDesc& L = net[li];
            const bool last = (li + 1 == n_layers);
            if (L.kind == KIND_GAP_LINEAR) {
                if (!last || !gap || cur.is_u8 || L.a.cout > 4) return fail(HELLO_ERR_UNSUPPORTED, "bad pooled head");
                if (!dry && n > 0) {
                    const long long threads = n * 32;
                    gap_linear_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
                        static_cast<const float*>(cur.base), n, cur.len, cur.ch, L.a.w, L.a.b, L.a.cout, gap->out,
                        gap->stride, gap->softmax);
                    if (!launched("gap_linear")) return false;
                }
                break;
            }
            if (last && !out) return fail(HELLO_ERR_UNSUPPORTED, "feature network without output buffer");
            if (L.kind == KIND_CONV) {
                const int lo = L.a.out_len(cur.len);
                int d = -1;
                float* y = last ? out : slab[d = grab()];
                if (!conv(cur, L.a, n, y, nullptr)) return false;
                if (cur_slab >= 0) used[cur_slab] = false;
                cur_slab = d;
                cur = view_cl(y, lo, L.a.cout);
            } else if (L.kind == KIND_MAXPOOL) {
                if (cur.is_u8 || cur.sc != 1 || cur.ch % 4) return fail(HELLO_ERR_UNSUPPORTED, "maxpool input");
                const int lo = (cur.len - L.a.k) / L.a.stride + 1;
                int d = -1;
                float* y = last ? out : slab[d = grab()];
                if (!dry && n > 0) {
                    const long long total = n * lo * (cur.ch / 4);
                    maxpool_kernel<<<grid_for(total, 256), 256, 0, st>>>(
                        static_cast<const float4*>(cur.base), reinterpret_cast<float4*>(y), n, cur.len, lo, cur.ch / 4,
                        L.a.k, L.a.stride);
                    if (!launched("maxpool")) return false;
                }
                if (cur_slab >= 0) used[cur_slab] = false;
                cur_slab = d;
                cur = view_cl(y, lo, cur.ch);
            } else if (L.kind == KIND_RES) {
                if (cur.is_u8 || cur.sc != 1 || cur.sl != cur.ch)
                    return fail(HELLO_ERR_UNSUPPORTED, "residual block on a strided input");
                const int lo = L.a.out_len(cur.len);
                const int ts = grab();
                if (!conv(cur, L.a, n, slab[ts], nullptr)) return false;
                const float* resid = static_cast<const float*>(cur.base);
                int ss = -1;
                if (L.has_shortcut) {
                    ss = grab();
                    if (!conv(cur, L.s, n, slab[ss], nullptr)) return false;
                    resid = slab[ss];
                } else if (L.a.cin != L.b.cout || lo != cur.len) {
                    return fail(HELLO_ERR_UNSUPPORTED, "identity shortcut with a shape change");
                }
                int d = -1;
                float* y = last ? out : slab[d = grab()];
                if (!conv(view_cl(slab[ts], lo, L.a.cout), L.b, n, y, resid)) return false;
                used[ts] = false;
                if (ss >= 0) used[ss] = false;
                if (cur_slab >= 0) used[cur_slab] = false;
                cur_slab = d;
                cur = view_cl(y, lo, L.b.cout);
            } else {
                return fail(HELLO_ERR_UNSUPPORTED, "unknown layer kind");
            }
        }
        arena->release(m);
        return status == HELLO_OK;
    }

    // fused tcgen05 head (headconv_tc.cuh): compressor / xattn (with in_s: 2a - s front end) / meta_convolver
    bool head(HeadConvTC* t, const float* in_a, const float* in_s, const int32_t* site_idx, long long n, float* out,
              long long out_stride, int softmax) {
        if (dry || n == 0) return true;
        h->launches++;
        return check(headconv_tc_launch(t, in_a, in_s, site_idx, n, out, out_stride, softmax, st), "headconv_tc");
    }

    // One head sub-network (compressor / xattn / meta_convolver) on n items [n, in_len, in_ch]; with in_s the operand is
    // 2*in_a - in_s[site_idx] (xattn_subtract.py:226-231).  Feature nets write `out`, pooled ones through `gap`.
    // Fused kernel when the layer table matches it, fused kernel + the added layers for an addendum model, else layer-wise.
    bool head_net(int id, const float* in_a, const float* in_s, const int32_t* site_idx, long long n, int in_len, int in_ch,
                  float* out, const GapOut* gap) {
        HeadConvTC* t = h->head[id];
        if (t && h->tail[id].empty())
            return gap ? head(t, in_a, in_s, site_idx, n, gap->out, gap->stride, gap->softmax)
                       : head(t, in_a, in_s, site_idx, n, out, 0, 0);
        const size_t m = arena->mark();
        bool ok;
        if (t) {
            float* mid = arena->allocf(n * (long long)t->out_len * t->out_ch);
            if (arena->overflow) return fail(HELLO_ERR_WORKSPACE, "workspace overflow");
            ok = head(t, in_a, in_s, site_idx, n, mid, 0, 0) &&
                 run_net(h->tail[id], view_cl(mid, t->out_len, t->out_ch), n, out, gap);
        } else if (in_s) {
            const long long elems = (long long)in_len * in_ch;
            float* x = arena->allocf(n * elems);
            if (arena->overflow) return fail(HELLO_ERR_WORKSPACE, "workspace overflow");
            ok = two_a_minus_s(in_a, in_s, site_idx, x, n, elems) && run_net(h->nets[id], view_cl(x, in_len, in_ch), n, out, gap);
        } else {
            ok = run_net(h->nets[id], view_cl(in_a, in_len, in_ch), n, out, gap);
        }
        arena->release(m);
        return ok;
    }

    bool comb(CombConvTC* t, const float* in_a, const float* in_b, int stride, long long n, float* out) {
        if (dry || n == 0) return true;
        h->launches++;
        return check(combconv_tc_launch(t, in_a, in_b, stride, n, out, st), "combconv_tc");
    }

    bool segsum(const float* x, float* out, const int32_t* off, long long n_groups, int row_base, long long elems) {
        if (dry || n_groups == 0) return true;
        const int e4 = (int)(elems / 4);
        dim3 grid((unsigned)n_groups, (unsigned)((e4 + 127) / 128));
        segsum_kernel<<<grid, 128, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(out), off,
                                            row_base, e4);
        return launched("segsum");
    }

    bool two_a_minus_s(const float* allele, const float* site, const int32_t* site_idx, float* out, long long na,
                       long long elems) {
        if (dry || na == 0) return true;
        const int e4 = (int)(elems / 4);
        two_a_minus_s_kernel<<<grid_for(na * e4, 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(allele), reinterpret_cast<const float4*>(site), site_idx,
            reinterpret_cast<float4*>(out), na, e4);
        return launched("two_a_minus_s");
    }

    bool add2(const float* a, const float* b, float* out, long long elems) {
        if (dry || elems == 0) return true;
        add2_kernel<<<grid_for(elems / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a),
                                                            reinterpret_cast<const float4*>(b),
                                                            reinterpret_cast<float4*>(out), elems / 4);
        return launched("add2");
    }

    bool concat2(const float* a, const float* b, float* out, long long rows, int ca, int cb) {
        if (dry || rows == 0) return true;
        concat2_kernel<<<grid_for(rows * (ca + cb) / 4, 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out),
            rows, ca / 4, cb / 4);
        return launched("concat2");
    }
};

struct Chunk {
    long long s0, s1, a0, a1, r0[2], r1[2];
    long long ns() const { return s1 - s0; }
    long long na() const { return a1 - a0; }
    long long nr(int t) const { return r1[t] - r0[t]; }
};

// One chunk of whole sites through the model. With run.dry only the arena high-water mark is computed.
bool forward_chunk(Runner& run, const Chunk& ck, const hello_batch* in, const hello_result* out) {
    hello_moe* h = run.h;
    const hello_cfg& cfg = h->cfg;
    Arena& ar = *run.arena;
    const bool dry = run.dry;
    const long long na = ck.na(), ns = ck.ns();
    const long long read_e = (long long)h->read_len * h->read_ch;
    const long long comp_e = (long long)h->comp_len * h->comp_ch;
    const int L = cfg.feature_length;
    const long long A_total = dry ? 0 : in->n_alleles;
    const size_t chunk_mark = ar.mark();

    int32_t* site_idx = static_cast<int32_t*>(ar.alloc(size_t(na) * sizeof(int32_t)));
    if (!dry && ns > 0) {
        site_index_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, run.st>>>(in->d_site_allele_off + ck.s0, (int)ns,
                                                                            (int)ck.a0, site_idx);
        if (!run.launched("site_index")) return false;
    }

    float* c_t[2] = {nullptr, nullptr};
    float* s_t[2] = {nullptr, nullptr};
    for (int t = 0; t < cfg.n_tech; ++t) {
        const long long nr = ck.nr(t);
        const int C = cfg.read_channels[t];
        c_t[t] = ar.allocf(na * comp_e);
        s_t[t] = ar.allocf(ns * comp_e);
        const size_t m1 = ar.mark();
        float* a_feat = ar.allocf(na * read_e);
        const size_t m2 = ar.mark();
        // the fused kernel sums the reads of every allele itself (no per-read maps in HBM); the per-layer path needs them
        const bool rc_tail = h->tc[t] && !h->tail[NET_RC0 + t].empty();
        // A handful of sites (the strict per-site call, a scoring-server batch): the fused allele sum makes a CTA own whole
        // alleles, so a 30-read allele is four work items in a row on ONE SM while the others idle.  Below this many reads
        // the kernel writes per-read maps instead (every CTA gets one item) and segsum_kernel adds them: the same additions
        // in the same order (test_huge_allele_and_partition_invariance), a quarter of the latency.
        const bool rc_small = h->tc[t] && !rc_tail && h->small_call[t];   // (per call, not per chunk: the chunk planner
                                                                          // relies on bigger chunks needing more workspace)
        float* r_feat = (h->tc[t] && !rc_tail && !rc_small) ? nullptr : ar.allocf(nr * read_e);
        float* r_pre = rc_tail ? ar.allocf(nr * read_e) : nullptr;      // fused original layers -> added layers
        if (ar.overflow) return run.fail(HELLO_ERR_WORKSPACE, "workspace overflow");
        // read convolver (architectures/read_convolver.py) on uint8 rows
        const uint8_t* reads = dry ? nullptr : in->d_reads[t] + (size_t)ck.r0[t] * L * C;
        cudaEvent_t ev_stop = nullptr;
        // Profiling brackets at most MAX_PROFILE_REGIONS regions between two hello_moe_profile_collect calls: a caller that
        // enables profiling and never collects does not grow the event pool without bound.
        if (!dry && h->profile && nr > 0 && h->ev_used + 2 <= 2 * MAX_PROFILE_REGIONS) {
            if (h->ev_used + 2 > h->ev_pool.size()) {
                cudaEvent_t a = nullptr, b = nullptr;
                if (!run.check(cudaEventCreate(&a), "cudaEventCreate") || !run.check(cudaEventCreate(&b), "cudaEventCreate"))
                    return false;
                h->ev_pool.push_back(a); h->ev_pool.push_back(b);
            }
            if (!run.check(cudaEventRecord(h->ev_pool[h->ev_used], run.st), "cudaEventRecord")) return false;
            ev_stop = h->ev_pool[h->ev_used + 1];
            h->ev_used += 2;
        }
        if (rc_tail) {
            if (!dry && nr > 0) {
                cudaError_t e = readconv_tc_launch(h->tc[t], reads, nr, in->input_layout, r_pre, run.st, nullptr, -1, nullptr,
                                                   nullptr, 0, 0);
                h->launches++;
                if (!run.check(e, "readconv_tc")) return false;
            }
            if (!run.run_net(h->tail[NET_RC0 + t], view_cl(r_pre, h->read_len, h->read_ch), nr, r_feat, nullptr)) return false;
            if (ev_stop && !run.check(cudaEventRecord(ev_stop, run.st), "cudaEventRecord")) return false;
            if (!run.segsum(r_feat, a_feat, dry ? nullptr : in->d_allele_read_off[t] + ck.a0, na, (int)ck.r0[t], read_e))
                return false;
        } else if (rc_small) {
            if (!dry && nr > 0) {
                cudaError_t e = readconv_tc_launch(h->tc[t], reads, nr, in->input_layout, r_feat, run.st, nullptr, -1, nullptr,
                                                   nullptr, 0, 0);
                h->launches++;
                if (!run.check(e, "readconv_tc")) return false;
            }
            if (ev_stop && !run.check(cudaEventRecord(ev_stop, run.st), "cudaEventRecord")) return false;
            if (!run.segsum(r_feat, a_feat, dry ? nullptr : in->d_allele_read_off[t] + ck.a0, na, (int)ck.r0[t], read_e))
                return false;
        } else if (h->tc[t]) {
            if (!dry && nr > 0) {
                // read convolver + reads -> alleles (reduceSlots, :163) in one kernel
                cudaError_t e = readconv_tc_launch(h->tc[t], reads, nr, in->input_layout, nullptr, run.st, nullptr, -1, a_feat,
                                                   in->d_allele_read_off[t] + ck.a0, na, (int)ck.r0[t]);
                h->launches++;
                if (!run.check(e, "readconv_tc")) return false;
            }
            if (ev_stop && !run.check(cudaEventRecord(ev_stop, run.st), "cudaEventRecord")) return false;
        } else {
            ActView v;
            v.base = reads; v.len = L; v.ch = C; v.is_u8 = true; v.sn = (long long)L * C;
            if (!dry && in->input_layout == HELLO_LAYOUT_RCL) { v.sc = L; v.sl = 1; } else { v.sc = 1; v.sl = C; }
            if (!run.run_net(h->nets[NET_RC0 + t], v, nr, r_feat, nullptr)) return false;
            if (ev_stop && !run.check(cudaEventRecord(ev_stop, run.st), "cudaEventRecord")) return false;
            // reads -> alleles (reduceSlots, :163)
            if (!run.segsum(r_feat, a_feat, dry ? nullptr : in->d_allele_read_off[t] + ck.a0, na, (int)ck.r0[t], read_e))
                return false;
        }
        ar.release(m2);
        // compressor (:125)
        if (!run.head_net(NET_CMP0 + t, a_feat, nullptr, nullptr, na, h->read_len, h->read_ch, c_t[t], nullptr)) return false;
        ar.release(m1);
        // alleles -> sites on the compressed features (:142-147)
        if (!run.segsum(c_t[t], s_t[t], dry ? nullptr : in->d_site_allele_off + ck.s0, ns, (int)ck.a0, comp_e))
            return false;
        if (cfg.xattn_present[t]) {
            Runner::GapOut g{dry ? nullptr : out->d_logits + (long long)t * A_total + ck.a0, 1, 0};
            if (!run.head_net(NET_X0 + t, c_t[t], s_t[t], site_idx, na, h->comp_len, h->comp_ch, nullptr, &g)) return false;
        }
    }

    bool meta_done = false;
    if (cfg.n_tech == 2 && cfg.xattn_present[2]) {
        // combiner path (:193-219): allele- and site-level fusion of the two technologies, then xattn2
        const int cc = h->comp_ch;
        float* c2 = ar.allocf(na * comp_e);
        float* s2 = ar.allocf(ns * comp_e);
        size_t m = ar.mark();
        if (cfg.has_combiners == HELLO_COMBINE_SUM) {
            // legacy wiring (MoEMergedAdvanced, useAdditive, no ConvCombiners; MixtureOfExpertsAdvanced.py:408-436): the hybrid
            // allele feature is the sum of the two technologies' and the hybrid site frame is reduceSlots of THAT
            if (!run.add2(c_t[0], c_t[1], c2, na * comp_e)) return false;
            if (!run.segsum(c2, s2, dry ? nullptr : in->d_site_allele_off + ck.s0, ns, (int)ck.a0, comp_e)) return false;
        } else if (h->comb[0] && h->comb[1]) {
            // the kernel reads the two technologies' tensors as the two K-halves: no concat buffer
            if (!run.comb(h->comb[0], c_t[0], c_t[1], cc, na, c2)) return false;
            if (!run.comb(h->comb[1], s_t[0], s_t[1], cc, ns, s2)) return false;
        } else {
            float* cat = ar.allocf(na * comp_e * 2);
            if (ar.overflow) return run.fail(HELLO_ERR_WORKSPACE, "workspace overflow");
            if (!run.concat2(c_t[0], c_t[1], cat, na * h->comp_len, cc, cc)) return false;
            if (!run.run_net(h->nets[NET_CB0], view_cl(cat, h->comp_len, 2 * cc), na, c2, nullptr)) return false;
            ar.release(m);
            cat = ar.allocf(ns * comp_e * 2);
            if (ar.overflow) return run.fail(HELLO_ERR_WORKSPACE, "workspace overflow");
            if (!run.concat2(s_t[0], s_t[1], cat, ns * h->comp_len, cc, cc)) return false;
            if (!run.run_net(h->nets[NET_CB1], view_cl(cat, h->comp_len, 2 * cc), ns, s2, nullptr)) return false;
            ar.release(m);
        }
        {
            Runner::GapOut g{dry ? nullptr : out->d_logits + 2LL * A_total + ck.a0, 1, 0};
            if (!run.head_net(NET_X2, c2, s2, site_idx, na, h->comp_len, cc, nullptr, &g)) return false;
        }
        if (cfg.meta_kind == HELLO_META_SITE) {
            Runner::GapOut gm{dry ? nullptr : out->d_meta + 3 * ck.s0, 3, 1};
            if (!run.head_net(NET_META, s2, nullptr, nullptr, ns, h->comp_len, cc, nullptr, &gm)) return false;
            meta_done = true;
        }
    }
    if (cfg.meta_kind == HELLO_META_REF) {
        ActView v;
        v.base = dry ? nullptr : in->d_ref_onehot + (size_t)ck.s0 * L * 5;
        v.sn = (long long)L * 5; v.sl = 5; v.sc = 1; v.len = L; v.ch = 5; v.is_u8 = false;
        Runner::GapOut gm{dry ? nullptr : out->d_meta + 3 * ck.s0, 3, 1};
        if (!run.run_net(h->nets[NET_META], v, ns, nullptr, &gm)) return false;
        meta_done = true;
    }
    if (!dry && ns > 0) {
        if (!meta_done) {
            fill_meta_default_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, run.st>>>(out->d_meta + 3 * ck.s0, ns);
            if (!run.launched("fill_meta")) return false;
        }
        PosteriorArgs pa;
        // wrapper semantics (:530-538): with a meta network all three rows go through the sigmoid (a missing
        // expert is a zero logit -> 0.5); without one the single returned head is expert 0 and the others are 0.
        const bool has_meta = cfg.meta_kind != HELLO_META_NONE;
        int head = 0;
        if (!has_meta) head = cfg.xattn_present[0] ? 0 : (cfg.xattn_present[2] ? 2 : 1);
        pa.logits = out->d_logits + (long long)head * A_total;
        pa.logit_stride = A_total;
        pa.expert_mask = has_meta ? 7 : 1;
        pa.meta = out->d_meta;
        pa.site_off = in->d_site_allele_off;
        pa.allele_rank = in->d_allele_rank;
        pa.pair_off = reinterpret_cast<const long long*>(in->d_pair_off);
        pa.pair_total = run.pair_total;
        pa.pair_prob = out->d_pair_prob;
        pa.pair_mix64 = out->d_pair_mix64;
        pa.best_pair = out->d_best_pair;
        pa.best_prob = out->d_best_prob;
        pa.call_pair = out->d_call_pair;
        pa.call_qual = out->d_call_qual;
        pa.best_expert = out->d_best_expert;
        pa.s_begin = ck.s0; pa.s_end = ck.s1;
        posterior_kernel<<<(unsigned)((ns * 32 + 255) / 256), 256, 0, run.st>>>(pa);
        if (!run.launched("posterior")) return false;
    }
    ar.release(chunk_mark);
    return run.status == HELLO_OK;
}

size_t dry_bytes(hello_moe* h, long long nr0, long long nr1, long long na, long long ns) {
    Arena ar; ar.dry = true;
    Runner run{h, &ar, nullptr, true};
    Chunk ck{};
    ck.s0 = 0; ck.s1 = ns; ck.a0 = 0; ck.a1 = na; ck.r0[0] = 0; ck.r1[0] = nr0; ck.r0[1] = 0; ck.r1[1] = nr1;
    std::string saved = h->err;
    forward_chunk(run, ck, nullptr, nullptr);
    h->err = saved;
    return ar.high + 256;
}

bool parse_blob(hello_moe* h, const void* blob, size_t nbytes, std::string& err) {
    const uint8_t* p = static_cast<const uint8_t*>(blob);
    if (nbytes < 128 || std::memcmp(p, "HELLOB2\0", 8) != 0) { err = "bad blob magic"; return false; }
    uint32_t version, n_slots; uint64_t rec_off, n_rec, data_off, n_floats;
    std::memcpy(&version, p + 8, 4); std::memcpy(&n_slots, p + 12, 4);
    std::memcpy(&rec_off, p + 16, 8); std::memcpy(&n_rec, p + 24, 8);
    std::memcpy(&data_off, p + 32, 8); std::memcpy(&n_floats, p + 40, 8);
    if (version != 1 || n_slots != N_NETS) { err = "unsupported blob version"; return false; }
    if (rec_off + n_rec * 128 > nbytes || data_off + n_floats * 4 > nbytes || data_off % 16) {
        err = "blob truncated"; return false;
    }
    uint32_t first[N_NETS], count[N_NETS];
    std::memcpy(first, p + 48, sizeof(first)); std::memcpy(count, p + 48 + sizeof(first), sizeof(count));
    if (cudaMalloc(&h->d_weights, std::max<size_t>(n_floats, 4) * 4) != cudaSuccess) { err = "cudaMalloc(weights) failed"; return false; }
    if (cudaMemcpy(h->d_weights, p + data_off, n_floats * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        err = "cudaMemcpy(weights) failed"; return false;
    }
    h->n_floats = n_floats;
    h->h_weights.resize(n_floats);
    std::memcpy(h->h_weights.data(), p + data_off, n_floats * 4);
    // `used`: the record's convolution slot is live (conv_a always for conv / residual / pooled-linear records, conv_b for
    // residual records, conv_s for residual records with a convolution shortcut).  Live slots must describe a real layer
    // whose weights [k*cin, cout] and bias [cout] lie inside the data section.
    auto conv_from = [&](const int32_t* r, ConvDesc* c, bool used) -> bool {
        c->cin = r[0]; c->cout = r[1]; c->k = r[2]; c->stride = r[3]; c->pad = r[4]; c->relu = r[5];
        if (r[6] < 0 || r[7] < 0 || (uint64_t)r[6] > n_floats || (uint64_t)r[7] > n_floats) return false;
        if (used) {
            if (r[0] <= 0 || r[1] <= 0 || r[2] <= 0 || r[3] <= 0 || r[4] < 0 || r[5] < ACT_NONE || r[5] > ACT_SOFTPLUS) return false;
            const uint64_t wn = (uint64_t)r[2] * (uint64_t)r[0] * (uint64_t)r[1];
            if ((uint64_t)r[6] + wn > n_floats || (uint64_t)r[7] + (uint64_t)r[1] > n_floats) return false;
        }
        c->w = h->d_weights + r[6]; c->b = h->d_weights + r[7];
        return true;
    };
    for (int n = 0; n < N_NETS; ++n) {
        if ((uint64_t)first[n] + count[n] > n_rec) { err = "bad net table"; return false; }
        for (uint32_t i = 0; i < count[n]; ++i) {
            int32_t r[32];
            std::memcpy(r, p + rec_off + (size_t)(first[n] + i) * 128, 128);
            LayerDesc L{};
            L.kind = r[0]; L.has_shortcut = r[1];
            if (L.kind < 0 || L.kind > KIND_GAP_LINEAR) { err = "unknown layer kind"; return false; }
            const bool res = L.kind == KIND_RES;
            if (L.kind == KIND_MAXPOOL) {
                if (r[4] <= 0 || r[5] <= 0) { err = "max-pool record with k or stride <= 0"; return false; }
            }
            if (!conv_from(r + 2, &L.a, L.kind != KIND_MAXPOOL) || !conv_from(r + 10, &L.b, res) ||
                !conv_from(r + 18, &L.s, res && L.has_shortcut != 0)) {
                err = "layer record out of range (cin / cout / k / stride must be positive, pad non-negative, activation 0..2, weights and "
                      "bias inside the data section)";
                return false;
            }
            h->nets[n].push_back(L);
        }
    }
    return true;
}

}  // namespace

extern "C" {

int hello_moe_abi_version(void) { return HELLO_MOE_ABI_VERSION; }

int hello_moe_create(const void* blob, size_t nbytes, const hello_cfg* cfg, int device, hello_moe** out) {
    if (!blob || !cfg || !out || cfg->struct_size != (int32_t)sizeof(hello_cfg)) {
        g_create_error = "hello_moe_create: bad arguments (struct_size mismatch?)";
        return HELLO_ERR_ARG;
    }
    *out = nullptr;
    if (cfg->n_tech < 1 || cfg->n_tech > 2 || cfg->feature_length <= 0) {
        g_create_error = "hello_moe_create: n_tech must be 1 or 2"; return HELLO_ERR_ARG;
    }
    if (cfg->precision != HELLO_PREC_FP32 && cfg->precision != HELLO_PREC_BF16X3 && cfg->precision != HELLO_PREC_BF16) {
        g_create_error = "hello_moe_create: precision must be one of HELLO_PREC_FP32 / HELLO_PREC_BF16X3 / HELLO_PREC_BF16";
        return HELLO_ERR_ARG;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return HELLO_ERR_CUDA;
    }
    hello_moe* h = new hello_moe();
    h->cfg = *cfg; h->device = device;
    std::string err;
    if (!parse_blob(h, blob, nbytes, err)) {
        g_create_error = "hello_moe_create: " + err;
        hello_moe_destroy(h);
        return HELLO_ERR_BLOB;
    }
    // shapes of the feature tensors and a consistency check of the wiring against the layer tables
    bool ok = true;
    for (int t = 0; t < cfg->n_tech && ok; ++t) {
        int l, c;
        ok = !h->nets[NET_RC0 + t].empty() && !h->nets[NET_CMP0 + t].empty() &&
             net_out_shape(h->nets[NET_RC0 + t], cfg->feature_length, cfg->read_channels[t], &l, &c);
        if (ok && t == 0) { h->read_len = l; h->read_ch = c; }
        ok = ok && l == h->read_len && c == h->read_ch;
        int l2, c2;
        ok = ok && net_out_shape(h->nets[NET_CMP0 + t], h->read_len, h->read_ch, &l2, &c2);
        if (ok && t == 0) { h->comp_len = l2; h->comp_ch = c2; }
        ok = ok && l2 == h->comp_len && c2 == h->comp_ch;
        ok = ok && (h->read_len * h->read_ch) % 4 == 0 && (h->comp_len * h->comp_ch) % 4 == 0 && h->comp_ch % 4 == 0;
    }
    for (int e3 = 0; e3 < 3 && ok; ++e3) {
        if (!cfg->xattn_present[e3]) continue;
        int l, c;
        ok = net_out_shape(h->nets[NET_X0 + e3], h->comp_len, h->comp_ch, &l, &c) && l == 1 && c == 1;
        if (e3 < 2) ok = ok && e3 < cfg->n_tech;
    }
    if (ok && cfg->xattn_present[2]) {
        int l, c;
        ok = cfg->n_tech == 2 && (cfg->has_combiners == HELLO_COMBINE_CONV || cfg->has_combiners == HELLO_COMBINE_SUM);
        for (int k = 0; k < 2 && ok && cfg->has_combiners == HELLO_COMBINE_CONV; ++k)
            ok = net_out_shape(h->nets[NET_CB0 + k], h->comp_len, 2 * h->comp_ch, &l, &c) && l == h->comp_len &&
                 c == h->comp_ch;
    }
    if (ok && cfg->meta_kind == HELLO_META_SITE) {
        int l, c;
        ok = cfg->xattn_present[2] && net_out_shape(h->nets[NET_META], h->comp_len, h->comp_ch, &l, &c) && c == 3;
    } else if (ok && cfg->meta_kind == HELLO_META_REF) {
        int l, c;
        ok = net_out_shape(h->nets[NET_META], cfg->feature_length, 5, &l, &c) && c == 3;
    }
    if (ok && cfg->meta_kind != HELLO_META_NONE) ok = cfg->n_tech == 2 && (cfg->xattn_present[0] && cfg->xattn_present[1]);
    if (ok && cfg->meta_kind == HELLO_META_NONE)
        ok = (cfg->xattn_present[0] + cfg->xattn_present[1] + cfg->xattn_present[2]) == 1 &&
             (cfg->n_tech == 1 ? cfg->xattn_present[0] : cfg->xattn_present[2]);
    if (!ok) {
        g_create_error = "hello_moe_create: wiring in hello_cfg does not match the layer tables of the blob";
        hello_moe_destroy(h);
        return HELLO_ERR_UNSUPPORTED;
    }
    if (cfg->precision != HELLO_PREC_FP32) {
        // Fused tcgen05 kernels where the sub-network has the architecture they are specialised for; every other
        // convolution (meta_convolver_ref, the 2x-wide models) goes through the generic tensor-core layer kernel.
        std::string terr;
        for (int t = 0; t < cfg->n_tech; ++t) {
            const std::vector<LayerDesc>& net = h->nets[NET_RC0 + t];
            size_t fl = 0;
            h->tc[t] = readconv_tc_create(net, h->d_weights, h->h_weights.data(), cfg->read_channels[t], cfg->feature_length,
                                          cfg->precision, terr, &fl);
            if (h->tc[t]) h->tail[NET_RC0 + t].assign(net.begin() + fl, net.end());
        }
        std::vector<int> want;
        for (int t = 0; t < cfg->n_tech; ++t) want.push_back(NET_CMP0 + t);
        for (int e3 = 0; e3 < 3; ++e3) if (cfg->xattn_present[e3]) want.push_back(NET_X0 + e3);
        if (cfg->meta_kind == HELLO_META_SITE) want.push_back(NET_META);
        for (int id : want) {
            const int in_len = (id == NET_CMP0 || id == NET_CMP1) ? h->read_len : h->comp_len;
            const std::vector<LayerDesc>& net = h->nets[id];
            size_t fl = 0;
            h->head[id] = headconv_tc_create(net, in_len, h->d_weights, h->h_weights.data(), cfg->precision, terr, &fl);
            if (h->head[id]) h->tail[id].assign(net.begin() + fl, net.end());
        }
        if (cfg->has_combiners == HELLO_COMBINE_CONV && cfg->xattn_present[2] && h->comp_len == cc::L) {
            for (int k = 0; k < 2; ++k)
                h->comb[k] = combconv_tc_create(h->nets[NET_CB0 + k], h->d_weights, h->h_weights.data(), cfg->precision, terr);
            if (!h->comb[0] || !h->comb[1]) {
                for (int k = 0; k < 2; ++k) { combconv_tc_destroy(h->comb[k]); h->comb[k] = nullptr; }
            }
        }
        h->layer_tc = convlayer_tc_create(cfg->precision, terr);
        bool ok_tc = h->layer_tc != nullptr;
        for (int n = 0; n < N_NETS && ok_tc; ++n) {
            const bool fused = (n <= NET_RC1 && h->tc[n - NET_RC0]) || h->head[n] ||
                               ((n == NET_CB0 || n == NET_CB1) && h->comb[n - NET_CB0]);
            for (const LayerDesc& L : fused ? h->tail[n] : h->nets[n]) {
                if (L.kind == KIND_CONV) ok_tc = convlayer_tc_add(h->layer_tc, L.a, h->d_weights, h->h_weights.data(), terr);
                if (L.kind == KIND_RES) {
                    ok_tc = convlayer_tc_add(h->layer_tc, L.a, h->d_weights, h->h_weights.data(), terr) &&
                            convlayer_tc_add(h->layer_tc, L.b, h->d_weights, h->h_weights.data(), terr, true) &&
                            (!L.has_shortcut || convlayer_tc_add(h->layer_tc, L.s, h->d_weights, h->h_weights.data(), terr));
                }
                if (!ok_tc) break;
            }
        }
        if (!ok_tc) {
            g_create_error = "hello_moe_create: tensor-core layers: " + terr;
            hello_moe_destroy(h);
            return HELLO_ERR_CUDA;
        }
    }
    *out = h;
    return HELLO_OK;
}

void hello_moe_destroy(hello_moe* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    for (int t = 0; t < 2; ++t) readconv_tc_destroy(h->tc[t]);
    for (int n = 0; n < N_NETS; ++n) headconv_tc_destroy(h->head[n]);
    for (int k = 0; k < 2; ++k) combconv_tc_destroy(h->comb[k]);
    convlayer_tc_destroy(h->layer_tc);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->d_weights) cudaFree(h->d_weights);
    delete h;
}

const char* hello_moe_last_error(const hello_moe* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

size_t hello_moe_workspace_bytes(const hello_moe* h, int64_t n_reads0, int64_t n_reads1, int64_t n_alleles,
                                 int64_t n_sites) {
    if (!h) return 0;
    hello_moe* hm = const_cast<hello_moe*>(h);
    hm->small_call[0] = n_reads0 <= SMALL_BATCH_READS;
    hm->small_call[1] = n_reads1 <= SMALL_BATCH_READS;
    return dry_bytes(hm, n_reads0, n_reads1, n_alleles, n_sites);
}

int hello_moe_profile_enable(hello_moe* h, int on) {
    if (!h) return HELLO_ERR_ARG;
    h->profile = on != 0;
    if (!on) h->ev_used = 0;
    return HELLO_OK;
}

int hello_moe_profile_collect(hello_moe* h, double* ms_read_conv, int64_t* n_regions) {
    if (!h || !ms_read_conv || !n_regions) return HELLO_ERR_ARG;
    double total = 0.0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
        cudaError_t e = cudaEventSynchronize(h->ev_pool[i + 1]);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]);
        if (e != cudaSuccess) { h->err = std::string("profile_collect: ") + cudaGetErrorString(e); return HELLO_ERR_CUDA; }
        total += ms;
    }
    *ms_read_conv = total;
    *n_regions = (int64_t)(h->ev_used / 2);
    h->ev_used = 0;
    return HELLO_OK;
}

int64_t hello_moe_launch_count(const hello_moe* h) { return h ? h->launches : 0; }

// Sites [sb, se) of the batch; n_pairs < 0 = count the genotype pairs of the whole batch here.
static int forward_impl(hello_moe* h, const hello_batch* in, const hello_result* out, long long sb, long long se,
                        long long n_pairs, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!h) return HELLO_ERR_ARG;
    h->err.clear();
    if (!in || !out || !d_workspace) { h->err = "null argument"; return HELLO_ERR_ARG; }
    const hello_cfg& cfg = h->cfg;
    const long long S = in->n_sites, A = in->n_alleles;
    if (S < 0 || A < S) { h->err = "need n_alleles >= n_sites >= 0"; return HELLO_ERR_ARG; }
    if (sb < 0 || se > S || sb > se) { h->err = "site range outside the batch"; return HELLO_ERR_ARG; }
    if (S == 0 || sb == se) return HELLO_OK;
    if (!in->h_site_allele_off || !in->d_site_allele_off || !in->d_pair_off || !out->d_logits || !out->d_meta ||
        !out->d_pair_prob || !out->d_best_pair || !out->d_best_prob) {
        h->err = "missing required buffer"; return HELLO_ERR_ARG;
    }
    if (in->h_site_allele_off[0] != 0 || in->h_site_allele_off[S] != A) {
        h->err = "site_allele_off must start at 0 and end at n_alleles"; return HELLO_ERR_ARG;
    }
    for (int t = 0; t < cfg.n_tech; ++t) {
        if (!in->d_reads[t] || !in->d_allele_read_off[t] || !in->h_allele_read_off[t]) {
            h->err = "missing read tensor / CSR for a technology"; return HELLO_ERR_ARG;
        }
        if (in->h_allele_read_off[t][0] != 0 || in->h_allele_read_off[t][A] != in->n_reads[t]) {
            h->err = "allele_read_off must start at 0 and end at n_reads"; return HELLO_ERR_ARG;
        }
    }
    if (in->input_layout != HELLO_LAYOUT_RCL && in->input_layout != HELLO_LAYOUT_RLC) {
        h->err = "bad input_layout"; return HELLO_ERR_ARG;
    }
    if (cfg.meta_kind == HELLO_META_REF && !in->d_ref_onehot) { h->err = "reference segment required"; return HELLO_ERR_ARG; }
    // Intentional deviation, stricter than the reference: reduceSlots (python/MixtureOfExpertsAdvanced.py:23-34) takes
    // results[cumsum(slots) - 1] minus the previous selection, so an empty slot after the first yields an all-zero sum
    // (and an empty FIRST slot indexes row -1, i.e. wraps to the last row).  The reference's own callers never produce
    // one -- an allele without support in a technology contributes one all-zero row
    // (AlleleSearcherLiteFiltered.cpp:1037-1043, AlleleSearcherLite.py:245-247) -- so an empty slot here means a broken
    // CSR and is reported as HELLO_ERR_ARG instead of being scored.  O(alleles of the range) on the host per call.
    for (int t = 0; t < cfg.n_tech; ++t) {
        const int32_t* off = in->h_allele_read_off[t];
        for (long long al = in->h_site_allele_off[sb]; al < in->h_site_allele_off[se]; ++al)
            if (off[al + 1] <= off[al]) { h->err = "every allele needs at least one read row per technology"; return HELLO_ERR_ARG; }
    }
    for (int t = 0; t < 2; ++t)
        h->small_call[t] = t < cfg.n_tech && (long long)in->h_allele_read_off[t][in->h_site_allele_off[se]] -
                                                 in->h_allele_read_off[t][in->h_site_allele_off[sb]] <= SMALL_BATCH_READS;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return HELLO_ERR_CUDA; }
    // absent experts read as zero logits (torch.zeros_like, :244)
    {
        const long long ab = in->h_site_allele_off[sb], ae = in->h_site_allele_off[se];
        for (int ex = 0; ex < 3 && e == cudaSuccess; ++ex)
            e = cudaMemsetAsync(out->d_logits + (size_t)ex * A + ab, 0, size_t(ae - ab) * sizeof(float), st);
    }
    if (e != cudaSuccess) { h->err = std::string("cudaMemsetAsync: ") + cudaGetErrorString(e); return HELLO_ERR_CUDA; }

    Arena ar;
    ar.base = static_cast<char*>(d_workspace);
    ar.cap = workspace_bytes;
    Runner run{h, &ar, st, false};
    for (long long s = (n_pairs < 0 ? 0 : sb); s < (n_pairs < 0 ? S : se); ++s) {
        const long long n = in->h_site_allele_off[s + 1] - in->h_site_allele_off[s];
        if (n < 1) { h->err = "every site needs at least one allele"; return HELLO_ERR_ARG; }
        run.pair_total += n * (n + 1) / 2;
    }
    if (n_pairs >= 0) run.pair_total = n_pairs;
    auto chunk_of = [&](long long s0, long long s1) {
        Chunk ck{};
        ck.s0 = s0; ck.s1 = s1;
        ck.a0 = in->h_site_allele_off[s0]; ck.a1 = in->h_site_allele_off[s1];
        for (int t = 0; t < cfg.n_tech; ++t) {
            ck.r0[t] = in->h_allele_read_off[t][ck.a0];
            ck.r1[t] = in->h_allele_read_off[t][ck.a1];
        }
        return ck;
    };
    auto fits = [&](const Chunk& ck) {
        return dry_bytes(h, ck.nr(0), ck.nr(1), ck.na(), ck.ns()) <= workspace_bytes;
    };
    long long s0 = sb;
    while (s0 < se) {
        long long cap = se - s0;
        if (cfg.max_chunk_sites > 0) cap = std::min<long long>(cap, cfg.max_chunk_sites);
        // largest chunk of whole sites that fits: gallop then bisect on the dry-run byte count
        long long lo = 1, hi = 1;
        if (!fits(chunk_of(s0, s0 + 1))) { h->err = "workspace too small for one site"; return HELLO_ERR_WORKSPACE; }
        while (hi < cap && fits(chunk_of(s0, s0 + std::min(cap, hi * 2)))) hi = std::min(cap, hi * 2);
        lo = hi; hi = std::min(cap, hi * 2);
        while (lo < hi) {
            const long long mid = (lo + hi + 1) / 2;
            if (fits(chunk_of(s0, s0 + mid))) lo = mid; else hi = mid - 1;
        }
        const Chunk ck = chunk_of(s0, s0 + lo);
        ar.top = 0;
        if (!forward_chunk(run, ck, in, out)) return run.status != HELLO_OK ? run.status : HELLO_ERR_CUDA;
        s0 += lo;
    }
    return HELLO_OK;
}

int hello_moe_forward(hello_moe* h, const hello_batch* in, const hello_result* out, void* d_workspace,
                      size_t workspace_bytes, void* stream) {
    return forward_impl(h, in, out, 0, in ? in->n_sites : 0, -1, d_workspace, workspace_bytes, stream);
}

int hello_moe_forward_range(hello_moe* h, const hello_batch* in, const hello_result* out, int64_t site_begin,
                            int64_t site_end, int64_t n_pairs, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (n_pairs < 0) { if (h) h->err = "n_pairs must be the batch's total pair count"; return HELLO_ERR_ARG; }
    return forward_impl(h, in, out, site_begin, site_end, n_pairs, d_workspace, workspace_bytes, stream);
}

int hello_moe_run_net(hello_moe* h, int net_id, const void* d_in, int64_t n_items, int32_t lin,
                      int32_t input_layout, float* d_out, int32_t* out_c, int32_t* out_l, void* d_workspace,
                      size_t workspace_bytes, void* stream) {
    if (!h) return HELLO_ERR_ARG;
    h->err.clear();
    if (net_id < 0 || net_id >= N_NETS || h->nets[net_id].empty() || !d_in || !d_out || n_items < 0) {
        h->err = "bad net id or buffers"; return HELLO_ERR_ARG;
    }
    const std::vector<LayerDesc>& net = h->nets[net_id];
    const LayerDesc& first = net.front();
    const int cin = first.a.cin;
    ActView v;
    const bool is_read = net_id == NET_RC0 || net_id == NET_RC1;
    if (is_read) {
        v.base = d_in; v.len = lin; v.ch = cin; v.is_u8 = true; v.sn = (long long)lin * cin;
        if (input_layout == HELLO_LAYOUT_RCL) { v.sc = lin; v.sl = 1; } else { v.sc = 1; v.sl = cin; }
    } else {
        v = view_cl(static_cast<const float*>(d_in), lin, cin);
    }
    int lo, co;
    if (!net_out_shape(net, lin, cin, &lo, &co)) { h->err = "shape mismatch"; return HELLO_ERR_ARG; }
    if (out_c) *out_c = co;
    if (out_l) *out_l = lo;
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return HELLO_ERR_CUDA; }
    Arena ar;
    ar.base = static_cast<char*>(d_workspace);
    ar.cap = workspace_bytes;
    Runner run{h, &ar, static_cast<cudaStream_t>(stream), false};
    const bool gap = net.back().kind == KIND_GAP_LINEAR;
    Runner::GapOut g{d_out, co, 0};
    if (is_read && h->tc[net_id - NET_RC0]) {
        const std::vector<LayerDesc>& tail = h->tail[net_id];
        float* pre = tail.empty() ? d_out : ar.allocf(n_items * (long long)lo * co);
        if (ar.overflow) { h->err = "workspace overflow"; return HELLO_ERR_WORKSPACE; }
        if (n_items > 0) {
            e = readconv_tc_launch(h->tc[net_id - NET_RC0], static_cast<const uint8_t*>(d_in), n_items,
                                   input_layout, pre, run.st);
            h->launches++;
            if (!run.check(e, "readconv_tc")) return run.status;
        }
        if (!tail.empty()) run.run_net(tail, view_cl(pre, lo, co), n_items, d_out, nullptr);
        return run.status;
    }
    if (h->head[net_id]) {
        run.head_net(net_id, static_cast<const float*>(d_in), nullptr, nullptr, n_items, lin, cin, gap ? nullptr : d_out,
                     gap ? &g : nullptr);
        return run.status;
    }
    if ((net_id == NET_CB0 || net_id == NET_CB1) && h->comb[net_id - NET_CB0]) {
        const float* x = static_cast<const float*>(d_in);       // [n, 18, 256]: the two halves of every row
        run.comb(h->comb[net_id - NET_CB0], x, x + cc::C_HALF, 2 * cc::C_HALF, n_items, d_out);
        return run.status;
    }
    run.run_net(net, v, n_items, gap ? nullptr : d_out, gap ? &g : nullptr);
    return run.status;
}

int hello_moe_readconv_debug(hello_moe* h, int tech, const uint8_t* d_reads, int64_t n_reads, int32_t input_layout,
                             int32_t phase, float* d_out, float* d_dbg, void* stream) {
    if (!h) return HELLO_ERR_ARG;
    h->err.clear();
    if (tech < 0 || tech > 1 || !h->tc[tech]) { h->err = "no tensor-core read convolver for this technology"; return HELLO_ERR_UNSUPPORTED; }
    if (!d_reads || !d_out || n_reads < 0) { h->err = "bad buffers"; return HELLO_ERR_ARG; }
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess)
        e = readconv_tc_launch(h->tc[tech], d_reads, n_reads, input_layout, d_out, static_cast<cudaStream_t>(stream),
                               d_dbg, d_dbg ? phase : -1);
    h->launches++;
    if (e != cudaSuccess) { h->err = std::string("readconv_tc: ") + cudaGetErrorString(e); return HELLO_ERR_CUDA; }
    return HELLO_OK;
}

int hello_moe_headconv_debug(hello_moe* h, int net_id, const float* d_in, int64_t n_items, int32_t phase, float* d_out,
                             float* d_dbg, void* stream) {
    if (!h) return HELLO_ERR_ARG;
    h->err.clear();
    if (!d_in || !d_out || n_items < 0) { h->err = "bad buffers"; return HELLO_ERR_ARG; }
    if ((net_id == NET_CB0 || net_id == NET_CB1) && h->comb[net_id - NET_CB0]) {
        cudaError_t ec = cudaSetDevice(h->device);
        if (ec == cudaSuccess)
            ec = combconv_tc_launch(h->comb[net_id - NET_CB0], d_in, d_in + cc::C_HALF, 2 * cc::C_HALF, n_items, d_out,
                                    static_cast<cudaStream_t>(stream), d_dbg, d_dbg ? phase : -1);
        h->launches++;
        if (ec != cudaSuccess) { h->err = std::string("combconv_tc: ") + cudaGetErrorString(ec); return HELLO_ERR_CUDA; }
        return HELLO_OK;
    }
    if (net_id < 0 || net_id >= N_NETS || !h->head[net_id]) { h->err = "no tensor-core head for this network"; return HELLO_ERR_UNSUPPORTED; }
    HeadConvTC* t = h->head[net_id];
    cudaError_t e = cudaSetDevice(h->device);
    if (e == cudaSuccess)
        e = headconv_tc_launch(t, d_in, nullptr, nullptr, n_items, d_out, t->prm.n_out > 0 ? t->prm.n_out : 0, 0,
                               static_cast<cudaStream_t>(stream), d_dbg, d_dbg ? phase : -1);
    h->launches++;
    if (e != cudaSuccess) { h->err = std::string("headconv_tc: ") + cudaGetErrorString(e); return HELLO_ERR_CUDA; }
    return HELLO_OK;
}

// ------------------------------------------------------------------------------------------- hello_encode.h
static thread_local std::string g_encode_error;

const char* hello_encode_last_error(void) { return g_encode_error.c_str(); }

int hello_encode_reads(const hello_encode_batch* b, uint8_t* d_out, void* stream) {
    g_encode_error.clear();
    if (!b || !d_out) { g_encode_error = "null argument"; return HELLO_ERR_ARG; }
    if (b->n_rows < 0 || b->feature_length < 1 || b->feature_length > enc::MAX_L || (b->channels != 6 && b->channels != 7)) {
        g_encode_error = "need n_rows >= 0, 1 <= feature_length <= 160, channels 6 or 7"; return HELLO_ERR_ARG;
    }
    if (b->n_rows == 0) return HELLO_OK;
    if (!b->d_row_read || !b->d_row_site || !b->d_read_off || !b->d_bases || !b->d_quals || !b->d_cigar_off || !b->d_cigars ||
        !b->d_ref_start || !b->d_mapq || !b->d_orientation || (b->channels == 7 && !b->d_hp) || !b->d_ref_off ||
        !b->d_reference || !b->d_window_start || !b->d_assembly_start || !b->d_assembly_stop) {
        g_encode_error = "missing buffer"; return HELLO_ERR_ARG;
    }
    static const enc::Luts luts = enc::make_luts();
    const long long rows_per_block = (long long)enc::WARPS * enc::ROWS_PER_WARP;
    const long long blocks = (b->n_rows + rows_per_block - 1) / rows_per_block;
    if (blocks > 0x7fffffffLL) { g_encode_error = "too many rows for one launch"; return HELLO_ERR_ARG; }
    enc::encode_reads_kernel<<<(unsigned)blocks, enc::WARPS * 32, 0, static_cast<cudaStream_t>(stream)>>>(*b, luts, d_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_encode_error = std::string("encode_reads: ") + cudaGetErrorString(e); return HELLO_ERR_CUDA; }
    return HELLO_OK;
}

}  // extern "C"
