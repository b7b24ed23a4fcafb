// Whole-network forward / backward of desed_task.nnet.CRNN (CRNN.py:221-306) as a stream-ordered kernel sequence over a
// caller-provided plan (include/sedk.h: sedk_crnn_plan).  No allocation, no host synchronisation: capturable in a CUDA graph.
//
//   forward : [scaler + specaug + conv0] -> per layer {conv3x3 (+BN stats) -> bn_finalize -> BN+GLU+dropout+pool}
//             -> [embedding fusion] -> per GRU layer {2 input GEMMs -> persistent bidirectional recurrence}
//             -> dropout -> heads (sigmoid / class-softmax attention pooling)
//   backward: heads -> dropout -> GRU BPTT (+ weight/input GEMMs) -> [fusion] -> per layer {BN+GLU+pool bwd -> BN bwd
//             apply -> conv wgrad -> conv dgrad}
// Dropout stream ids: conv layer i -> i, embedding concat -> 100, post-RNN -> 200 (masks are regenerated, never stored).
#include "kernels.h"

namespace sedk {
namespace {

constexpr uint64_t STREAM_EMB = 100, STREAM_RNN = 200;

// Fork / join helper: kernels that are off the dependency chain (weight packing in forward; the weight-gradient GEMMs of
// the GRU and of the convolutions in backward) run on a library-owned side stream so that they overlap the chain instead
// of extending it.  Event record / wait pairs are capturable: inside a CUDA graph they become parallel branches.
// Disabled while per-launcher profiling is on (the event brackets would time overlapped kernels) or with option
// "side_stream" = 0.
struct Side {
    cudaStream_t s = nullptr;
    cudaEvent_t ev[16];
    int next = 0;
    bool ok = false;
    Side() {
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return;
        for (int i = 0; i < 16; i++)
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) return;
        ok = true;
    }
    cudaEvent_t event() { return ev[next++ & 15]; }
};
Side& side() {
    static thread_local Side x;
    return x;
}
// returns the side stream after making it wait for everything enqueued on `main` so far, or `main` itself when disabled
struct Fork {
    cudaStream_t main, side_s;
    bool forked = false;
    explicit Fork(cudaStream_t m) : main(m), side_s(m) {}
    int begin() {
        if (forked || profiling_on() || get_option("side_stream", 1) == 0 || !side().ok) return SEDK_OK;
        cudaEvent_t e = side().event();
        SEDK_CUDA(cudaEventRecord(e, main));
        SEDK_CUDA(cudaStreamWaitEvent(side().s, e, 0));
        side_s = side().s;
        forked = true;
        return SEDK_OK;
    }
    // side work issued so far must see what `main` has done up to now (a new dependency edge main -> side)
    int sync_side_to_main() {
        if (!forked) return SEDK_OK;
        cudaEvent_t e = side().event();
        SEDK_CUDA(cudaEventRecord(e, main));
        SEDK_CUDA(cudaStreamWaitEvent(side_s, e, 0));
        return SEDK_OK;
    }
    // main waits for all side work issued so far
    int join() {
        if (!forked) return SEDK_OK;
        cudaEvent_t e = side().event();
        SEDK_CUDA(cudaEventRecord(e, side_s));
        SEDK_CUDA(cudaStreamWaitEvent(main, e, 0));
        return SEDK_OK;
    }
};

// the tcgen05 BN+GLU path needs its two workspaces and covers C = 128, pooling (1, 2), TF32 mode; the backward of a layer
// must take the same path as its forward (different dropout-mask mapping, lin saved), so both sides ask this one function
bool use_glu_tc5(const sedk_crnn_plan* p, const sedk_conv_layer& L) {
    return p->activation == 0 && L.glu_pack != nullptr && L.lin != nullptr && bnglu_tc5_supports(L.T, L.F, L.cout, L.pt, L.pf, p->precision);
}

// store-free first block (layer0.cu): shipped geometry, GLU, and the caller gave the sums workspace
bool use_l0_fused(const sedk_crnn_plan* p) {
    const sedk_conv_layer& L = p->conv[0];
    return p->l0_sums != nullptr && p->x0 != nullptr && p->activation == 0 && L.cin == 1 &&
           l0_fused_supports(p->B, L.T, L.F, L.cout, L.pt, L.pf);
}

template <class... P>
bool all_aligned16(P... ptrs) {
    return ((((reinterpret_cast<uintptr_t>(ptrs)) & 15) == 0) && ...);
}

int validate(const sedk_crnn_plan* p, bool backward) {
    SEDK_REQUIRE(p != nullptr, "crnn: null plan");
    SEDK_REQUIRE(p->B > 0 && p->n_conv >= 1 && p->n_conv <= SEDK_MAX_CONV, "crnn: bad B / n_conv");
    SEDK_REQUIRE(p->n_gru >= 1 && p->n_gru <= SEDK_MAX_GRU_LAYERS, "crnn: bad n_gru");
    SEDK_REQUIRE(p->activation >= 0 && p->activation <= 3, "crnn: unknown activation %d", p->activation);
    SEDK_REQUIRE(p->x && p->strong && p->weak && p->sof && p->hsum, "crnn: missing input / output buffers");
    SEDK_REQUIRE(p->conv[0].cin == 1, "crnn: the first conv layer must have one input channel (n_in_channel=1)");
    SEDK_REQUIRE(p->conv[0].T == p->n_frames && p->conv[0].F == p->n_mels, "crnn: layer-0 geometry mismatch");
    for (int i = 0; i < p->n_conv; i++) {
        const sedk_conv_layer& L = p->conv[i];
        SEDK_REQUIRE(L.w && L.b && L.gamma && L.beta && L.running_mean && L.running_var,
                     "crnn: conv layer %d has null parameters", i);
        SEDK_REQUIRE(p->activation >= 2 || (L.glu_w && L.glu_b), "crnn: conv layer %d has no gate parameters", i);
        SEDK_REQUIRE(L.z && L.out && L.stats && L.bn, "crnn: conv layer %d has null workspace", i);
        if (i > 0) {
            SEDK_REQUIRE(L.wpack, "crnn: conv layer %d needs wpack", i);
            SEDK_REQUIRE(L.cin == p->conv[i - 1].cout && L.T == p->conv[i - 1].T / p->conv[i - 1].pt &&
                             L.F == p->conv[i - 1].F / p->conv[i - 1].pf,
                         "crnn: conv layer %d geometry does not chain", i);
        }
        if (backward) {
            SEDK_REQUIRE(L.gw && L.gb && L.ggamma && L.gbeta && L.gy && L.gout,
                         "crnn: conv layer %d has null gradient buffers", i);
            SEDK_REQUIRE(p->activation >= 2 || (L.gglu_w && L.gglu_b), "crnn: conv layer %d has no gate gradient buffers", i);
            SEDK_REQUIRE(i == 0 || L.gwpack, "crnn: conv layer %d needs gwpack", i);
        }
    }
    const sedk_conv_layer& last = p->conv[p->n_conv - 1];
    SEDK_REQUIRE(last.F / last.pf == 1, "crnn: the CNN must pool the mel axis down to 1 (got %d)", last.F / last.pf);
    if (p->l0_sums != nullptr && p->zero_fwd != nullptr && p->zero_fwd_bytes > 0) {
        const char* lo = (const char*)p->zero_fwd;
        const char* q = (const char*)p->l0_sums;
        SEDK_REQUIRE(q >= lo && q + SEDK_L0_SUMS * sizeof(double) <= lo + p->zero_fwd_bytes,
                     "crnn: l0_sums must lie inside zero_fwd when zero_fwd is used");
    }
    if (backward) {
        SEDK_REQUIRE(p->training, "crnn backward: the forward pass must have run with training=1");
        SEDK_REQUIRE(p->x0, "crnn backward: x0 workspace missing");
        SEDK_REQUIRE(p->gdense_w && p->gdense_b && p->gsoft_w && p->gsoft_b && p->grnn_drop,
                     "crnn backward: head gradient buffers missing");
    }
    return SEDK_OK;
}

}  // namespace
}  // namespace sedk

using namespace sedk;

extern "C" int sedk_crnn_forward(const sedk_crnn_plan* p, void* stream) {
    int rc = validate(p, false);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int B = p->B;
    const float pdrop = p->training ? p->dropout_p : 0.f;
    // batch statistics only in a training forward whose BatchNorm is not frozen (freeze_bn / eval-mode autograd)
    const int bn_train = (p->training && !p->bn_eval) ? 1 : 0;
    const bool zf = p->zero_fwd != nullptr && p->zero_fwd_bytes > 0;
    if (zf && p->training) SEDK_CUDA(cudaMemsetAsync(p->zero_fwd, 0, (size_t)p->zero_fwd_bytes, s));
    // ---------------- weight packs of every layer: off the chain, overlapped with the first conv
    Fork fk(s);
    rc = fk.begin();
    if (rc) return rc;
    for (int i = 1; i < p->n_conv; i++) {
        const sedk_conv_layer& Lp = p->conv[i];
        rc = launch_pack_weights(Lp.w, Lp.wpack, Lp.cin, Lp.cout,
                                 conv_pair_mode(Lp.cin, Lp.cout, Lp.F, p->precision) ? 2 : (p->precision == 0 ? 1 : 0),
                                 fk.side_s);
        if (rc) return rc;
    }
    if (p->n_conv < 2) {
        rc = fk.join();
        if (rc) return rc;
    }
    // ---------------- CNN
    for (int i = 0; i < p->n_conv; i++) {
        const sedk_conv_layer& L = p->conv[i];
        const int C = L.cout;
        if (p->training && !zf) SEDK_CUDA(cudaMemsetAsync(L.stats, 0, 4 * C * sizeof(double), s));
        double* st = bn_train ? L.stats : nullptr;
        if (i == 0 && use_l0_fused(p)) {
            if (bn_train && !zf) SEDK_CUDA(cudaMemsetAsync(p->l0_sums, 0, SEDK_L0_SUMS * sizeof(double), s));
            rc = launch_l0_prep(p->x, p->x_sb, p->x_sm, p->x_st, p->minmax, p->scaler_eps,
                                p->training ? p->specaug : nullptr, L.w, L.b, p->x0, st, p->l0_sums, B, L.T, L.F, s);
            if (rc) return rc;
            rc = launch_bn_finalize(L.stats, L.gamma, L.beta, L.running_mean, L.running_var, L.num_batches, L.bn,
                                    (double)B * L.T * L.F, p->bn_eps, p->bn_momentum, bn_train, C, s);
            if (rc) return rc;
            rc = launch_l0_fwd(p->x0, L.w, L.b, L.bn, L.glu_w, L.glu_b, L.out, B, L.T, L.F, pdrop, p->seed, p->seed_dev,
                               (uint64_t)0, p->precision, s);
            if (rc) return rc;
            continue;
        }
        if (i == 0) {
            rc = launch_conv0_fwd(p->x, p->x_sb, p->x_sm, p->x_st, p->minmax, p->scaler_eps,
                                  p->training ? p->specaug : nullptr, L.w, L.b, p->training ? p->x0 : nullptr, L.z, st, B,
                                  L.T, L.F, C, s);
        } else {
            if (i == 1) {
                rc = fk.join();
                if (rc) return rc;
            }
            rc = launch_conv3x3_layer(p->conv[i - 1].out, L.wpack, 0, L.b, L.z, st, B, L.T, L.F, L.cin, C, p->precision, s);
        }
        if (rc) return rc;
        if (use_glu_tc5(p, L)) {
            rc = launch_glu_prep(L.stats, L.gamma, L.beta, L.running_mean, L.running_var, L.num_batches, L.bn, L.glu_w,
                                 L.glu_b, L.glu_pack, (double)B * L.T * L.F, p->bn_eps, p->bn_momentum, bn_train, C, s);
            if (rc) return rc;
            rc = launch_bnglu_tc5_fwd(L.z, L.bn, L.glu_pack, L.out, p->training ? L.lin : nullptr, B, L.T, L.F, C, L.pt, L.pf,
                                      pdrop, p->seed, p->seed_dev, (uint64_t)i, s);
            if (rc) return rc;
            continue;
        }
        rc = launch_bn_finalize(L.stats, L.gamma, L.beta, L.running_mean, L.running_var, L.num_batches, L.bn,
                                (double)B * L.T * L.F, p->bn_eps, p->bn_momentum, bn_train, C, s);
        if (rc) return rc;
        rc = launch_bnglu_pool_fwd(L.z, L.bn, L.glu_w, L.glu_b, L.out, B, L.T, L.F, C, L.pt, L.pf, pdrop, p->seed,
                                   p->seed_dev, (uint64_t)i, p->precision, p->activation, s);
        if (rc) return rc;
    }
    const sedk_conv_layer& last = p->conv[p->n_conv - 1];
    const int Tp = last.T / last.pt;
    const int nb = last.cout;
    const float* xr = last.out;      // [B, T', nb]
    int in_dim = nb;
    // ---------------- embedding fusion
    if (p->emb != nullptr) {
        SEDK_REQUIRE(p->cat_w && p->cat_b && p->cat_in && p->fused, "crnn: embedding fusion buffers missing");
        SEDK_REQUIRE(p->emb_mode == 0 || p->emb_mode == 1, "crnn: unknown emb_mode %d", p->emb_mode);
        rc = launch_emb_concat(xr, p->emb, p->training ? p->dropstep : nullptr, p->cat_in, B, Tp, nb, p->emb_dim,
                               p->emb_T, p->emb_mode, pdrop, p->seed, p->seed_dev, STREAM_EMB, s);
        if (rc) return rc;
        const int W = nb + p->emb_dim;
        // cat_tf (CRNN.py:294): [B T', nb + emb] x [nb, nb + emb]^T - on tcgen05 in the TF32 mode (K = 896 for BEATs embeddings)
        if (gemm_tc5_ok(B * Tp, nb, W, p->precision) && all_aligned16(p->cat_in, p->cat_w, p->cat_b, p->fused))
            rc = launch_gemm_tc5_nt1(p->cat_in, p->cat_w, p->cat_b, p->fused, B * Tp, nb, W, s);
        else
            rc = launch_gemm(0, 1, B * Tp, nb, W, 1.f, p->cat_in, W, p->cat_w, W, 0.f, p->fused, nb, p->cat_b, p->precision, s);
        if (rc) return rc;
        xr = p->fused;
    } else if (p->training && p->dropstep != nullptr) {
        // no embeddings (CRNN.py:295-301): dropstep span mask + dropout on the CNN output feed the GRU
        SEDK_REQUIRE(p->cat_in, "crnn: dropstep without embeddings needs the cat_in workspace [B, T', nb]");
        rc = launch_emb_concat(xr, nullptr, p->dropstep, p->cat_in, B, Tp, nb, 0, 0, 0, pdrop, p->seed, p->seed_dev,
                               STREAM_EMB, s);
        if (rc) return rc;
        xr = p->cat_in;
    }
    // ---------------- BiGRU
    for (int l = 0; l < p->n_gru; l++) {
        const sedk_gru_layer& G = p->gru[l];
        SEDK_REQUIRE(G.in_dim == in_dim, "crnn: GRU layer %d expects %d inputs, gets %d", l, G.in_dim, in_dim);
        const int H = G.hidden;
        for (int d = 0; d < 2; d++)
            SEDK_REQUIRE(G.w_ih[d] && G.w_hh[d] && G.b_ih[d] && G.b_hh[d] && G.gi[d], "crnn: GRU layer %d null", l);
        {
            const float* As[2] = {xr, xr};
            const float* Bs[2] = {G.w_ih[0], G.w_ih[1]};
            float* Cs[2] = {G.gi[0], G.gi[1]};
            const float* bs[2] = {G.b_ih[0], G.b_ih[1]};
            if (gemm_tc5_ok(B * Tp, 3 * H, in_dim, p->precision) && all_aligned16(xr, Bs[0], Bs[1], Cs[0], Cs[1], bs[0], bs[1]))
                rc = launch_gemm_tc5_nt2(xr, Bs, bs, Cs, B * Tp, 3 * H, in_dim, s);
            else
                rc = launch_gemm_batched(0, 1, B * Tp, 3 * H, in_dim, 1.f, As, in_dim, Bs, in_dim, 0.f, Cs, 3 * H, bs, 2,
                                         p->precision, s);
            if (rc) return rc;
        }
        SEDK_REQUIRE(G.out && (!p->training || (G.gates[0] && G.gates[1] && G.hprev[0] && G.hprev[1])),
                     "crnn: GRU layer %d workspace missing", l);
        rc = launch_gru_seq_fwd(G.gi, G.w_hh, G.b_hh, G.out, G.gates, G.hprev, B, Tp, H, p->training, s);
        if (rc) return rc;
        xr = G.out;
        in_dim = 2 * H;
    }
    // ---------------- dropout + heads
    const float* hx = xr;
    if (pdrop > 0.f) {
        SEDK_REQUIRE(p->rnn_drop, "crnn: rnn_drop workspace missing");
        rc = launch_dropout(xr, p->rnn_drop, (int64_t)B * Tp * in_dim, pdrop, p->seed, p->seed_dev, STREAM_RNN, s);
        if (rc) return rc;
        hx = p->rnn_drop;
    }
    return launch_heads_fwd(hx, p->dense_w, p->dense_b, p->soft_w, p->soft_b, p->classes_mask, p->strong, p->weak,
                            p->sof, p->hsum, B, Tp, in_dim, p->nclass, s);
}

// phases (bit mask): 1 = heads + BiGRU (+ embedding fusion), 4 = conv layers [3, n_conv) (98 % of the CNN parameters), 8 = conv
// layers [0, 3); 2 = 4 | 8 = the whole CNN, 3 / 15 = everything.  A call that does not run everything joins its side-stream
// work before it returns, so the gradients of the layers it covered are final (a data-parallel caller reduces that slice of
// the flat gradient while the next phase runs).
static int crnn_backward_impl(const sedk_crnn_plan* p, int phases, void* stream) {
    int rc = validate(p, true);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int B = p->B;
    const float pdrop = p->dropout_p;
    const sedk_conv_layer& last = p->conv[p->n_conv - 1];
    const int Tp = last.T / last.pt;
    const int nb = last.cout;
    const int BT = B * Tp;
    const int Hl = p->gru[p->n_gru - 1].hidden;
    const int D = 2 * Hl, C = p->nclass;
    Fork fk(s);
    // zb: every gradient buffer was cleared by one memset; zs: the backward halves of `stats` were cleared by the forward
    const bool zb = p->zero_bwd != nullptr && p->zero_bwd_bytes > 0;
    const bool zs = p->zero_fwd != nullptr && p->zero_fwd_bytes > 0;
    if (zb && (phases & 1)) SEDK_CUDA(cudaMemsetAsync(p->zero_bwd, 0, (size_t)p->zero_bwd_bytes, s));
    if (phases & 1) {
    // ---------------- heads
    if (!zb) {
        SEDK_CUDA(cudaMemsetAsync(p->gdense_w, 0, (size_t)C * D * sizeof(float), s));
        SEDK_CUDA(cudaMemsetAsync(p->gsoft_w, 0, (size_t)C * D * sizeof(float), s));
        SEDK_CUDA(cudaMemsetAsync(p->gdense_b, 0, (size_t)C * sizeof(float), s));
        SEDK_CUDA(cudaMemsetAsync(p->gsoft_b, 0, (size_t)C * sizeof(float), s));
    }
    const float* hx = pdrop > 0.f ? p->rnn_drop : p->gru[p->n_gru - 1].out;
    float* ghx = pdrop > 0.f ? p->grnn_drop : p->gru[p->n_gru - 1].gout;
    rc = launch_heads_bwd(hx, p->dense_w, p->soft_w, p->classes_mask, p->strong, p->hsum, p->sof, p->gstrong, p->gweak,
                          ghx, p->gdense_w, p->gdense_b, p->gsoft_w, p->gsoft_b, B, Tp, D, C, s);
    if (rc) return rc;
    if (pdrop > 0.f) {
        rc = launch_dropout(p->grnn_drop, p->gru[p->n_gru - 1].gout, (int64_t)BT * D, pdrop, p->seed, p->seed_dev, STREAM_RNN, s);
        if (rc) return rc;
    }
    // ---------------- BiGRU
    for (int l = p->n_gru - 1; l >= 0; l--) {
        const sedk_gru_layer& G = p->gru[l];
        const int H = G.hidden, in_dim = G.in_dim;
        SEDK_REQUIRE(G.gout && G.dghn[0] && G.dghn[1], "crnn backward: GRU layer %d gradient workspace missing", l);
        SEDK_REQUIRE(G.gb_ih[0] && G.gb_ih[1] && G.gb_hh[0] && G.gb_hh[1], "crnn backward: GRU bias grads null");
        rc = launch_gru_seq_bwd(G.gout, G.w_hh, G.gates, G.hprev, G.gi, G.dghn, G.gb_ih, G.gb_hh, B, Tp, H, zb ? 1 : 0, s);
        if (rc) return rc;
        const bool drop_only = p->emb == nullptr && p->dropstep != nullptr;      // CRNN.py:295-301
        const float* xin = l > 0 ? p->gru[l - 1].out : (p->emb ? p->fused : (drop_only ? p->cat_in : last.out));
        float* gin = l > 0 ? p->gru[l - 1].gout : ((p->emb || drop_only) ? p->gfused : last.gout);
        SEDK_REQUIRE(gin, "crnn backward: input-gradient buffer of GRU layer %d missing", l);
        // the weight-gradient GEMMs only feed the optimiser: side stream
        rc = l == p->n_gru - 1 ? fk.begin() : fk.sync_side_to_main();
        if (rc) return rc;
        cudaStream_t ss = fk.side_s;
        for (int d = 0; d < 2; d++) {
            SEDK_REQUIRE(G.gw_ih[d] && G.gw_hh[d] && G.gb_ih[d] && G.gb_hh[d], "crnn backward: GRU grads null");
            if (zb) continue;
            SEDK_CUDA(cudaMemsetAsync(G.gw_ih[d], 0, (size_t)3 * H * in_dim * sizeof(float), ss));
            SEDK_CUDA(cudaMemsetAsync(G.gw_hh[d], 0, (size_t)3 * H * H * sizeof(float), ss));
        }
        {
            // dW_ih = dgi^T x ; dW_hh rows [0,2H) from (dr, dz), rows [2H,3H) from d(hn) - both directions per launch
            const float* dgi[2] = {G.gi[0], G.gi[1]};
            const float* xs[2] = {xin, xin};
            float* gwih[2] = {G.gw_ih[0], G.gw_ih[1]};
            rc = launch_gemm_batched(1, 0, 3 * H, in_dim, BT, 1.f, dgi, 3 * H, xs, in_dim, 1.f, gwih, in_dim, nullptr, 2,
                                     p->precision, ss);
            if (rc) return rc;
            const float* hp[2] = {G.hprev[0], G.hprev[1]};
            float* gwhh[2] = {G.gw_hh[0], G.gw_hh[1]};
            rc = launch_gemm_batched(1, 0, 2 * H, H, BT, 1.f, dgi, 3 * H, hp, H, 1.f, gwhh, H, nullptr, 2, p->precision, ss);
            if (rc) return rc;
            const float* dhn[2] = {G.dghn[0], G.dghn[1]};
            float* gwhn[2] = {G.gw_hh[0] + (size_t)2 * H * H, G.gw_hh[1] + (size_t)2 * H * H};
            rc = launch_gemm_batched(1, 0, H, H, BT, 1.f, dhn, H, hp, H, 1.f, gwhn, H, nullptr, 2, p->precision, ss);
            if (rc) return rc;
        }
        if (gemm_tc5_ok(BT, in_dim, 3 * H, p->precision) && all_aligned16(G.gi[0], G.gi[1], G.w_ih[0], G.w_ih[1], gin)) {
            // dx = dgi_f W_ih,f + dgi_b W_ih,b : one tcgen05 launch, split-K over the two directions
            const float* da[2] = {G.gi[0], G.gi[1]};
            const float* wb[2] = {G.w_ih[0], G.w_ih[1]};
            rc = launch_gemm_tc5_nn_pair(da, wb, gin, BT, in_dim, 3 * H, s);
            if (rc) return rc;
        } else {
            for (int d = 0; d < 2; d++) {
                // dx (+)= dgi W_ih
                rc = launch_gemm(0, 0, BT, in_dim, 3 * H, 1.f, G.gi[d], 3 * H, G.w_ih[d], in_dim, d == 0 ? 0.f : 1.f, gin,
                                 in_dim, nullptr, p->precision, s);
                if (rc) return rc;
            }
        }
    }
    // ---------------- embedding fusion
    if (p->emb == nullptr && p->dropstep != nullptr) {
        SEDK_REQUIRE(p->gfused && last.gout, "crnn backward: dropstep gradient buffers missing");
        rc = launch_emb_concat_bwd(p->gfused, p->dropstep, last.gout, B, Tp, nb, 0, pdrop, p->seed, p->seed_dev, STREAM_EMB, s);
        if (rc) return rc;
    }
    if (p->emb != nullptr) {
        SEDK_REQUIRE(p->gcat_w && p->gcat_b && p->gfused && last.gout, "crnn backward: fusion gradient buffers missing");
        const int W = nb + p->emb_dim;
        if (!zb) SEDK_CUDA(cudaMemsetAsync(p->gcat_w, 0, (size_t)nb * W * sizeof(float), s));
        rc = launch_gemm(1, 0, nb, W, BT, 1.f, p->gfused, nb, p->cat_in, W, 1.f, p->gcat_w, W, nullptr, p->precision, s);
        if (rc) return rc;
        rc = launch_colsum(p->gfused, BT, nb, nb, p->gcat_b, 0, s);
        if (rc) return rc;
        // gcat = gfused cat_w, written over cat_in (no longer needed)
        rc = launch_gemm(0, 0, BT, W, nb, 1.f, p->gfused, nb, p->cat_w, W, 0.f, p->cat_in, W, nullptr, p->precision, s);
        if (rc) return rc;
        rc = launch_emb_concat_bwd(p->cat_in, p->dropstep, last.gout, B, Tp, nb, p->emb_dim, pdrop, p->seed, p->seed_dev, STREAM_EMB,
                                   s);
        if (rc) return rc;
    }
    }  // phase 1
    // split mode: the RNN-side weight-gradient GEMMs (side stream) must be complete when phase 1 returns; in the one-call
    // mode they keep overlapping the CNN backward and are joined at the very end
    if (!(phases & (4 | 8))) return fk.join();
    // ---------------- CNN
    constexpr int kSplit = 3;
    const int i_hi = (phases & 4) ? p->n_conv - 1 : (p->n_conv < kSplit ? p->n_conv : kSplit) - 1;
    const int i_lo = (phases & 8) ? 0 : kSplit;
    for (int i = i_hi; i >= i_lo; i--) {
        const sedk_conv_layer& L = p->conv[i];
        const int Cc = L.cout;
        const int64_t npix = (int64_t)B * L.T * L.F;
        if (!zs) SEDK_CUDA(cudaMemsetAsync(L.stats + 2 * Cc, 0, 2 * Cc * sizeof(double), s));
        if (!zb && L.gglu_b) SEDK_CUDA(cudaMemsetAsync(L.gglu_b, 0, (size_t)Cc * sizeof(float), s));
        if (i == 0 && use_l0_fused(p)) {
            if (!zb) {
                SEDK_CUDA(cudaMemsetAsync(L.gglu_w, 0, (size_t)Cc * Cc * sizeof(float), s));
                SEDK_CUDA(cudaMemsetAsync(L.gw, 0, (size_t)9 * Cc * sizeof(float), s));
            }
            // the GX block of l0_sums is still zero from the forward's clear (only this kernel adds to it)
            rc = launch_l0_bwd(p->x0, L.w, L.b, L.bn, L.glu_w, L.glu_b, L.gout, L.gglu_w, L.gglu_b, L.stats, p->l0_sums,
                               L.gw, L.gb, L.ggamma, L.gbeta, B, L.T, L.F, p->bn_eval ? 1 : 0, pdrop, p->seed,
                               p->seed_dev, (uint64_t)0, p->precision, s);
            if (rc) return rc;
            continue;
        }
        if (use_glu_tc5(p, L)) {
            rc = launch_bnglu_tc5_bwd(L.z, L.bn, L.glu_pack, L.gout, L.lin, L.gy, L.gglu_b, L.stats, B, L.T, L.F, Cc, L.pt,
                                      L.pf, pdrop, p->seed, p->seed_dev, (uint64_t)i, s);
            if (rc) return rc;
            // gate weight gradient = g_lin^T z over all pixels: off the chain
            rc = fk.forked ? fk.sync_side_to_main() : fk.begin();
            if (rc) return rc;
            if (!zb) SEDK_CUDA(cudaMemsetAsync(L.gglu_w, 0, (size_t)Cc * Cc * sizeof(float), fk.side_s));
            rc = launch_glu_wgrad_tc5(L.z, L.lin, L.bn, L.glu_pack, L.gglu_w, L.gglu_b, B, L.T, L.F, Cc, fk.side_s);
            if (rc) return rc;
        } else {
        if (!zb && L.gglu_w) SEDK_CUDA(cudaMemsetAsync(L.gglu_w, 0, (size_t)Cc * Cc * sizeof(float), s));
        rc = launch_bnglu_pool_bwd(L.z, L.bn, L.glu_w, L.glu_b, L.gout, L.gy, L.gglu_w, L.gglu_b, L.stats, B, L.T, L.F, Cc,
                                   L.pt, L.pf, pdrop, p->seed, p->seed_dev, (uint64_t)i, p->precision, p->activation, s);
        if (rc) return rc;
        }
        rc = launch_bn_bwd_apply(L.gy, L.z, L.bn, L.stats, L.ggamma, L.gbeta, L.gb, (double)npix, npix, Cc,
                                 p->bn_eval ? 1 : 0, s);
        if (rc) return rc;
        if (i > 0) {
            const sedk_conv_layer& P = p->conv[i - 1];
            SEDK_REQUIRE(P.gout, "crnn backward: gout of conv layer %d missing", i - 1);
            rc = fk.forked ? fk.sync_side_to_main() : fk.begin();
            if (rc) return rc;
            if (!zb) SEDK_CUDA(cudaMemsetAsync(L.gwpack, 0, (size_t)9 * Cc * L.cin * sizeof(float), fk.side_s));
            rc = launch_conv_wgrad(P.out, L.gy, L.gwpack, B, L.T, L.F, L.cin, Cc, p->precision, fk.side_s);
            if (rc) return rc;
            rc = launch_unpack_wgrad(L.gwpack, L.gw, L.cin, Cc, fk.side_s);
            if (rc) return rc;
            rc = launch_conv3x3_layer(L.gy, L.wpack, 1, nullptr, P.gout, nullptr, B, L.T, L.F, L.cin, Cc, p->precision, s);
            if (rc) return rc;
        } else {
            if (!zb) SEDK_CUDA(cudaMemsetAsync(L.gw, 0, (size_t)9 * Cc * sizeof(float), s));
            rc = launch_conv0_wgrad(p->x0, L.gy, L.gw, B, L.T, L.F, Cc, p->precision, s);
            if (rc) return rc;
        }
    }
    return fk.join();
}

extern "C" int sedk_crnn_backward(const sedk_crnn_plan* p, void* stream) { return crnn_backward_impl(p, 1 | 4 | 8, stream); }

extern "C" int sedk_crnn_backward_phase(const sedk_crnn_plan* p, int phases, void* stream) {
    if (phases < 1 || phases > 15) {
        sedk::set_error("sedk_crnn_backward_phase: phases is a bit mask in [1, 15] (got %d)", phases);
        return SEDK_ERR_INVALID;
    }
    if (phases & 2) phases = (phases & ~2) | 4 | 8;
    return crnn_backward_impl(p, phases, stream);
}
