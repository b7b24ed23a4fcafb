// Internal launcher prototypes shared by the libsedk translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace sedk {

// ---- conv.cu ------------------------------------------------------------------------------------------------
// wpack[0][tap][co][ci] = w[co][ci][tap];  wpack[1][tap][ci][co] = w[co][ci][8 - tap]   (tap = ky*3 + kx)
// round_tf32: round the packed weights to the nearest TF32 value (TF32 mode; the tcgen05 unit truncates its operands)
int launch_pack_weights(const float* w, float* wpack, int cin, int cout, int round_tf32, cudaStream_t s);
// gw[co][ci][tap] = gwpack[tap][co][ci]
int launch_unpack_wgrad(const float* gwpack, float* gw, int cin, int cout, cudaStream_t s);
// first layer (Cin = 1): fused instance-minmax scaler + specaugment mask + 3x3 stencil.
//   x: log-mel (b, m, t) strided; writes x0 [B,T,F] (the scaled/masked conv input, for backward; may be NULL),
//   z [B,T,F,cout] (+bias) and accumulates per-channel sum / sum^2 into stats[0..2*cout) when stats != NULL.
int launch_conv0_fwd(const float* x, int64_t sb, int64_t sm, int64_t st, const uint32_t* minmax, float scaler_eps,
                     const int32_t* specaug, const float* w, const float* bias, float* x0, float* z, double* stats,
                     int B, int T, int F, int cout, cudaStream_t s);
// gw[co][tap] += sum_pix gz[pix][co] * x0[pix + shift(tap)]     (gw must be zeroed by the caller)
int launch_conv0_wgrad(const float* x0, const float* gz, float* gw, int B, int T, int F, int cout, int precision,
                       cudaStream_t s);
// generic 3x3 / pad 1 / stride 1 conv on channels-last tensors with tensor-core MMA:
//   out[b,t,f,n] = bias[n] + sum_{tap,k} in[b,t+dy,f+dx,k] * wp[tap][n][k]; optional per-channel sum / sum^2.
// The same kernel computes dgrad when given the flipped/transposed pack (wpack[1]).
int launch_conv3x3(const float* in, const float* wp, const float* bias, float* out, double* stats, int B, int T, int F,
                   int cin, int cout, int precision, cudaStream_t s);
// Layer-level entry used by the network: picks the operand pack inside the layer's `wpack` workspace itself.
//   wpack_base: the buffer launch_pack_weights filled for the layer (cin_l -> cout_l); dgrad = 0: forward (cin_l -> cout_l),
//   dgrad = 1: data gradient (cout_l -> cin_l, flipped taps).  For 16-channel layers in the TF32 mode the buffer holds the
//   PAIRED-PIXEL packs (see launch_pack_weights) and the convolution runs on tcgen05 as a 2 cin -> 2 cout layer over F / 2.
int launch_conv3x3_layer(const float* in, const float* wpack_base, int dgrad, const float* bias, float* out, double* stats,
                         int B, int T, int F, int cin_l, int cout_l, int precision, cudaStream_t s);
bool conv_pair_mode(int cin_l, int cout_l, int F, int precision);
int conv_wpack_floats(int cin_l, int cout_l);   // size of the layer's wpack workspace
// tcgen05 / TMEM / TMA implicit-GEMM variant for 128 -> 128 channels (TF32); conv_tc5.cu
bool tc5_enabled();
void tc5_set(int on);
bool tc5_supports(int cin, int cout);
bool tc5_wgrad_supports(int cin, int cout);
int launch_conv_wgrad_tc5(const float* x, const float* gz, float* gwpack, int B, int T, int F, int cin, int cout,
                          cudaStream_t s);
int launch_tn_gemm_tc5_c128(const float* x, const float* g, float* out, int B, int T, int F, cudaStream_t s);
// cmod = cout, or cout / 2 in the paired-pixel mode (bias / BatchNorm statistics are indexed modulo cmod)
int launch_conv3x3_tc5(const float* in, const float* wp, const float* bias, float* out, double* stats, int B, int T,
                       int F, int cin, int cout, int cmod, cudaStream_t s);
// gwpack[tap][co][ci] += sum_pix gz[pix][co] * x[pix + shift(tap)][ci]   (gwpack zeroed by the caller)
int launch_conv_wgrad(const float* x, const float* gz, float* gwpack, int B, int T, int F, int cin, int cout,
                      int precision, cudaStream_t s);

// ---- bnglu.cu -----------------------------------------------------------------------------------------------
// bn[0]=scale, bn[1]=shift, bn[2]=mean, bn[3]=invstd from batch sums (training) or running stats (eval);
// training also updates running_mean / running_var (momentum, unbiased var) and num_batches.
int launch_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, int64_t* num_batches, float* bn, double count, float eps, float momentum,
                       int training, int C, cudaStream_t s);
// out = avgpool( dropout( act(y) ) ),  y = scale*z + shift;  act (CNN.py:81-88): 0 GLU (Wg y + bg) * sigmoid(y),
// 1 ContextGating y * sigmoid(Wg y + bg), 2 ReLU, 3 LeakyReLU(0.2) (glu_w / glu_b / their gradients NULL for 2, 3)
int launch_bnglu_pool_fwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, float* out, int B,
                          int T, int F, int C, int pt, int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                          uint64_t drop_stream, int precision, int act, cudaStream_t s);
// backward of the block above: writes gy (grad wrt y = BN output) for every pooled pixel, accumulates gglu_w, gglu_b
// and stats[2C..4C) = {sum gy, sum gy*zhat}.  gy must be pre-zeroed when the pooling drops rows/cols.
int launch_bnglu_pool_bwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, const float* gout,
                          float* gy, float* gglu_w, float* gglu_b, double* stats, int B, int T, int F, int C, int pt,
                          int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream, int precision,
                          int act, cudaStream_t s);
// bnglu_small.cu: register-resident warp-autonomous variants of the two kernels above for C in {16, 32} (pf == 2)
bool bnglu_small_supports(int B, int T, int F, int C, int pt, int pf);
int launch_bnglu_small_fwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, float* out, int B,
                           int T, int F, int C, int pt, int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                           uint64_t drop_stream, int precision, cudaStream_t s);
int launch_bnglu_small_bwd(const float* z, const float* bn, const float* glu_w, const float* glu_b, const float* gout,
                           float* gy, float* gglu_w, float* gglu_b, double* stats, int B, int T, int F, int C, int pt,
                           int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream,
                           int precision, cudaStream_t s);
// bnglu_tc5.cu: tcgen05 / TMEM / TMA BN+GLU(+dropout+pool (1,2)) for the 128-channel layers, TF32 mode.
//   launch_glu_prep = bn_finalize + gate-weight packs: pack = [W' = Wg*scale | WT = Wg^T | b' = bg + Wg shift]
bool bnglu_tc5_supports(int T, int F, int C, int pt, int pf, int precision);
int bnglu_tc5_pack_floats();     // size of the `glu_pack` workspace (floats)
int launch_glu_prep(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                    int64_t* num_batches, float* bn, const float* glu_w, const float* glu_b, float* pack, double count,
                    float eps, float momentum, int training, int C, cudaStream_t s);
int launch_bnglu_tc5_fwd(const float* z, const float* bn, const float* pack, float* out, float* lin_out, int B, int T, int F,
                         int C, int pt, int pf, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream,
                         cudaStream_t s);
int launch_bnglu_tc5_bwd(const float* z, const float* bn, const float* pack, const float* gout, float* lin_glin, float* gy,
                         float* gglu_b, double* stats, int B, int T, int F, int C, int pt, int pf, float drop_p, uint64_t seed,
                         const uint64_t* seed_dev, uint64_t drop_stream, cudaStream_t s);
int launch_glu_wgrad_tc5(const float* z, const float* g_lin, const float* bn, float* pack, float* gglu_w, const float* gglu_b,
                         int B, int T, int F, int C, cudaStream_t s);
// BN backward: gy -> gz in place; writes ggamma, gbeta and the conv-bias grad gb.  frozen = 0: batch-statistics BatchNorm
// (gb = 0: the bias cancels); frozen = 1: running-statistics BatchNorm, gz = scale * gy, gb = scale * sum gy
int launch_bn_bwd_apply(float* gy, const float* z, const float* bn, const double* stats, float* ggamma, float* gbeta,
                        float* gb, double count, int64_t n_pix, int C, int frozen, cudaStream_t s);

// ---- layer0.cu ---------------------------------------------------------------------------------------------
// First block (1 -> 16 filters, GLU, pooling (2, 2)) with the conv output never written to HBM: see the file header.
// sums = sedk_crnn_plan.l0_sums (304 doubles: GX [16][9], ZX [16][9], SX [16]), zero before a training forward.
bool l0_fused_supports(int B, int T, int F, int cout, int pt, int pf);
// x (strided log-mel) -> x0 [B,T,F] (scaler + SpecAugment); stats != NULL: also sum z / sum z^2 into stats[0..32) and
// ZX / SX into sums (batch-statistics forward)
int launch_l0_prep(const float* x, int64_t sb, int64_t sm, int64_t st, const uint32_t* minmax, float scaler_eps,
                   const int32_t* specaug, const float* w, const float* bias, float* x0, double* stats, double* sums,
                   int B, int T, int F, cudaStream_t s);
int launch_l0_fwd(const float* x0, const float* w, const float* bias, const float* bn, const float* glu_w,
                  const float* glu_b, float* out, int B, int T, int F, float drop_p, uint64_t seed,
                  const uint64_t* seed_dev, uint64_t drop_stream, int precision, cudaStream_t s);
// accumulates gglu_w / gglu_b / stats[32..64) / sums.GX, then writes gw (+=), gb, ggamma, gbeta in closed form
int launch_l0_bwd(const float* x0, const float* w, const float* bias, const float* bn, const float* glu_w,
                  const float* glu_b, const float* gout, float* gglu_w, float* gglu_b, double* stats, double* sums,
                  float* gw, float* gb, float* ggamma, float* gbeta, int B, int T, int F, int frozen, float drop_p,
                  uint64_t seed, const uint64_t* seed_dev, uint64_t drop_stream, int precision, cudaStream_t s);

// ---- gemm.cu ------------------------------------------------------------------------------------------------
int launch_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                int ldb, float beta, float* C, int ldc, const float* bias, int precision, cudaStream_t s);
// up to 4 problems of identical shape in one launch (blockIdx.z); pointer arrays live on the host
int launch_gemm_batched(int transA, int transB, int M, int N, int K, float alpha, const float* const* A, int lda,
                        const float* const* B, int ldb, float beta, float* const* C, int ldc, const float* const* bias,
                        int nprob, int precision, cudaStream_t s);
// gemm_tc5.cu: tcgen05 variants of the two GEMM shapes on the GRU's dependency chain (TF32 mode, K % 32 == 0, 16-byte
// aligned operands): C_d = A B_d^T + bias_d for d = 0, 1 in one launch, and C = A0 B0 + A1 B1 (split-K over two pairs)
bool gemm_tc5_ok(int M, int N, int K, int precision);
int launch_gemm_tc5_nt2(const float* A, const float* const B[2], const float* const bias[2], float* const C[2], int M, int N,
                        int K, cudaStream_t s);
int launch_gemm_tc5_nt1(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, cudaStream_t s);
int launch_gemm_tc5_nn_pair(const float* const A[2], const float* const B[2], float* C, int M, int N, int K, cudaStream_t s);
// column sums: out[n] (+)= sum_m A[m*lda + n]
int launch_colsum(const float* A, int M, int N, int lda, float* out, int accumulate, cudaStream_t s);

// ---- gru.cu -------------------------------------------------------------------------------------------------
// one direction-pair launch: gi[dir] [B,T,3H] (= x W_ih^T + b_ih), out [B,T,2H]; saves gates/hprev when training
int launch_gru_seq_fwd(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
                       float* const gates[2], float* const hprev[2], int B, int T, int H, int save, cudaStream_t s);
// BPTT: gout [B,T,2H] -> dgi[dir] [B,T,3H] (written over gi), dghn[dir] [B,T,H] and the bias gradients gb_ih / gb_hh [3H]
int launch_gru_seq_bwd(const float* gout, const float* const w_hh[2], const float* const gates[2],
                       const float* const hprev[2], float* const dgi[2], float* const dghn[2], float* const gb_ih[2],
                       float* const gb_hh[2], int B, int T, int H, int zeroed, cudaStream_t s);

// gru3.cu: third-generation H = 128 recurrence, one (row, direction) per CTA (variant 1: 8 warps, all weights in registers;
// variant 2: 16 warps)
int launch_gru_fwd_v3(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
                      float* const gates[2], float* const hprev[2], int B, int T, int save, int variant, cudaStream_t s);
int launch_gru_bwd_v3(const float* gout, const float* const w_hh[2], const float* const gates[2],
                      const float* const hprev[2], float* const dgi[2], float* const dghn[2], float* const gb_ih[2],
                      float* const gb_hh[2], int B, int T, int zeroed, int variant, cudaStream_t s);

// H = 192: 3-CTA cluster per (row, direction), h / dgh exchanged through DSMEM, one cluster barrier per step
int launch_gru_fwd_c3(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
                      float* const gates[2], float* const hprev[2], int B, int T, int save, cudaStream_t s);
int launch_gru_bwd_c3(const float* gout, const float* const w_hh[2], const float* const gates[2],
                      const float* const hprev[2], float* const dgi[2], float* const dghn[2], float* const gb_ih[2],
                      float* const gb_hh[2], int B, int T, int zeroed, cudaStream_t s);

// ---- heads.cu -----------------------------------------------------------------------------------------------
int launch_heads_fwd(const float* x, const float* dw, const float* db, const float* sw, const float* sb,
                     const uint8_t* cmask, float* strong, float* weak, float* sof, float* hsum, int B, int T, int D,
                     int C, cudaStream_t s);
int launch_heads_bwd(const float* x, const float* dw, const float* sw, const uint8_t* cmask, const float* strong,
                     const float* hsum, const float* sof, const float* gstrong, const float* gweak, float* gx,
                     float* gdw, float* gdb, float* gsw, float* gsb, int B, int T, int D, int C, cudaStream_t s);
// y = keep ? x/(1-p) : 0 (stream-indexed Philox mask); used forward and backward
int launch_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev,
                   uint64_t stream_id, cudaStream_t s);
// time-aggregation (mode 0: adaptive_avg_pool1d, 1: nearest-exact interpolate) + dropstep span masks + concat + dropout for
// the embedding fusion (CRNN.py:270-294) and its backward (x part only).  emb_dim = 0: the embedding-free branch
// (CRNN.py:295-301: dropstep mask + dropout on the CNN output)
int launch_emb_concat(const float* x, const float* emb, const int32_t* dropstep, float* cat, int B, int T, int nb,
                      int emb_dim, int emb_T, int mode, float p, uint64_t seed, const uint64_t* seed_dev, uint64_t stream_id,
                      cudaStream_t s);
int launch_emb_concat_bwd(const float* gcat, const int32_t* dropstep, float* gx, int B, int T, int nb, int emb_dim,
                          float p, uint64_t seed, const uint64_t* seed_dev, uint64_t stream_id, cudaStream_t s);

}  // namespace sedk
