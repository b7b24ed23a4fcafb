// Persistent bidirectional GRU recurrence (nn.GRU semantics, gate order r,z,n; desed_task/nnet/RNN.py:19-30).
//
// The input-side GEMMs (x W_ih^T + b_ih for all T steps) are hoisted out (gemm.cu).  What is left is strictly
// sequential: h_t = f(W_hh h_{t-1}, gi_t).  Batch rows are independent in the recurrence, so a CTA owns NB batch rows
// of one direction and keeps W_hh RESIDENT IN REGISTERS for all T steps: thread (row j, 64-wide k segment) holds 64
// weights; per step it does 64*NB FMAs against h (broadcast from shared memory), segment partials meet in shared
// memory, and H*NB "gate" threads finish the cell (exact fp32; tanhf/expf).  Two __syncthreads per step, no global
// traffic on the dependency chain except the prefetched gi row.  The backward kernel mirrors it with W_hh^T.
//
// H = 128 (2023 recipe) runs as a single CTA of 768 threads.  H = 192 (2024 recipe) does not fit one SM's register
// file, so the hidden units are split over a cluster of 3 CTAs (64 units each, 576 threads) which exchange the new h
// through distributed shared memory each step.
#include "kernels.h"
#include <cooperative_groups.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace sedk {
namespace {

constexpr int SEG = 64;

// fast, accurate-enough gate nonlinearities (MUFU.EX2 based: abs error ~1e-7, far inside the 1e-3 posterior budget)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// The register file cannot hold all of W_hh next to the loop state (49 152 of 65 536 registers at H = 128: the first
// version spilled inside the recurrence), so each thread keeps WR = 40 of its 64 weights in registers and the last 24 in
// shared memory, laid out [chunk][thread] so that a warp's 16-byte reads are contiguous.
// (with >= 768 threads per CTA; smaller CTAs - the cluster variants - keep all 64 in registers)
template <int NT>
struct WSplit {
    static constexpr int WR = NT >= 700 ? 40 : 64;      // weights per thread held in registers
    static constexpr int WS4 = (SEG - WR) / 4;          // float4 chunks per thread held in shared memory
};

// dot product of 64 weights (40 registers + 24 shared) with 64 shared-memory values; packed FP32 FMA (FFMA2, sm_100+)
template <int WR>
__device__ __forceinline__ float dot64_ffma2(const float2 (&w)[WR / 2], const float4* __restrict__ ws, int nthreads,
                                             const float* __restrict__ v) {
    constexpr int WS4 = (SEG - WR) / 4;
    const float4* v4 = reinterpret_cast<const float4*>(v);
    // four independent accumulator chains (the packed FMA has ~4-cycle dependent latency) and the shared-memory operands
    // of a whole group are fetched before they are consumed
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f), a3 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k8 = 0; k8 < WR / 8; k8++) {
        const float4 h0 = v4[2 * k8], h1 = v4[2 * k8 + 1];
        a0 = __ffma2_rn(w[4 * k8], make_float2(h0.x, h0.y), a0);
        a1 = __ffma2_rn(w[4 * k8 + 1], make_float2(h0.z, h0.w), a1);
        a2 = __ffma2_rn(w[4 * k8 + 2], make_float2(h1.x, h1.y), a2);
        a3 = __ffma2_rn(w[4 * k8 + 3], make_float2(h1.z, h1.w), a3);
    }
#pragma unroll
    for (int c = 0; c < WS4; c += 2) {
        const float4 h0 = v4[WR / 4 + c], h1 = v4[WR / 4 + c + 1];
        const float4 w0 = ws[c * nthreads], w1 = ws[(c + 1) * nthreads];
        a0 = __ffma2_rn(make_float2(w0.x, w0.y), make_float2(h0.x, h0.y), a0);
        a1 = __ffma2_rn(make_float2(w0.z, w0.w), make_float2(h0.z, h0.w), a1);
        a2 = __ffma2_rn(make_float2(w1.x, w1.y), make_float2(h1.x, h1.y), a2);
        a3 = __ffma2_rn(make_float2(w1.z, w1.w), make_float2(h1.z, h1.w), a3);
    }
    return ((a0.x + a0.y) + (a1.x + a1.y)) + ((a2.x + a2.y) + (a3.x + a3.y));
}

template <int H, int CS>
struct GruCfg {
    static constexpr int HU = H / CS;             // hidden units owned by one CTA
    static constexpr int R = 3 * HU;              // gate rows owned by one CTA
    static constexpr int SEGS = H / SEG;          // k segments (forward) per row
    static constexpr int NT_F = R * SEGS;         // forward threads
    static constexpr int JSEGS = 3 * H / SEG;     // j segments (backward) per column
    static constexpr int NT_B = HU * JSEGS;       // backward threads
    static_assert(H % SEG == 0 && H % CS == 0 && (HU % 32) == 0, "unsupported hidden size");
    static_assert(NT_F <= 1024 && NT_B <= 1024, "too many threads");
};
template <int H, int CS, int NB>
struct GruNbOk {
    static_assert(NB * GruCfg<H, CS>::HU <= GruCfg<H, CS>::NT_F && NB * GruCfg<H, CS>::HU <= GruCfg<H, CS>::NT_B,
                  "not enough threads for the gate phase");
    static constexpr bool ok = true;
};

// ----------------------------------------------------------------------------------------------------------------
template <int H, int CS, int NB>
__global__ void __launch_bounds__(GruCfg<H, CS>::NT_F, 1)
gru_fwd_kernel(const float* __restrict__ gi0, const float* __restrict__ gi1, const float* __restrict__ whh0,
               const float* __restrict__ whh1, const float* __restrict__ bhh0, const float* __restrict__ bhh1,
               float* __restrict__ out, float* __restrict__ gates0, float* __restrict__ gates1,
               float* __restrict__ hprev0, float* __restrict__ hprev1, int B, int T, int save) {
    using Cfg = GruCfg<H, CS>;
    constexpr int HU = Cfg::HU, R = Cfg::R, SEGS = Cfg::SEGS, NT = Cfg::NT_F;
    constexpr int WR = WSplit<NT>::WR, WS4 = WSplit<NT>::WS4;
    extern __shared__ __align__(16) float gru_smem[];
    float4* ws4 = reinterpret_cast<float4*>(gru_smem);                                  // [WS4][NT] float4
    float (*h_s)[NB][H] = reinterpret_cast<float (*)[NB][H]>(gru_smem + WS4 * NT * 4);  // [2][NB][H]
    float (*part)[NB][R] = reinterpret_cast<float (*)[NB][R]>(gru_smem + WS4 * NT * 4 + 2 * NB * H);   // [SEGS][NB][R]
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    int crank = 0;
    if (CS > 1) crank = (int)cg::this_cluster().block_rank();
    const int b0 = (blockIdx.x / CS) * NB;
    const int u0 = crank * HU;                              // first hidden unit owned by this CTA
    const float* gi = dir ? gi1 : gi0;
    const float* whh = dir ? whh1 : whh0;
    const float* bhh = dir ? bhh1 : bhh0;
    float* gates = dir ? gates1 : gates0;
    float* hprev = dir ? hprev1 : hprev0;

    const int seg = tid / R, row = tid - seg * R;           // row in [0, R): gate = row / HU, unit = row % HU
    const int gate = row / HU, unit = row - gate * HU;
    const int grow = gate * H + u0 + unit;                  // row of W_hh [3H, H]
    float2 w[WR / 2];
#pragma unroll
    for (int k = 0; k < WR / 2; k++)
        w[k] = *reinterpret_cast<const float2*>(whh + (size_t)grow * H + seg * SEG + 2 * k);
#pragma unroll
    for (int c = 0; c < WS4; c++)
        ws4[c * NT + tid] = *reinterpret_cast<const float4*>(whh + (size_t)grow * H + seg * SEG + WR + 4 * c);

    for (int i = tid; i < 2 * NB * H; i += NT) (&h_s[0][0][0])[i] = 0.f;
    const bool is_gate = tid < NB * HU;
    const int gb = tid / HU, gu = tid - gb * HU;            // batch row / local unit of a gate thread
    const int bglob = b0 + gb;
    const bool active = is_gate && (bglob < B);
    float bhr = 0.f, bhz = 0.f, bhn = 0.f, hval = 0.f;
    if (is_gate) {
        bhr = bhh[u0 + gu];
        bhz = bhh[H + u0 + gu];
        bhn = bhh[2 * H + u0 + gu];
    }
    if (CS > 1) cg::this_cluster().sync(); else __syncthreads();

    int cur = 0;
    // the gi row of step s+1 is fetched while step s computes (the only global read on the dependency chain)
    float gir = 0.f, giz = 0.f, gin = 0.f;
    if (active) {
        const float* gp = gi + ((size_t)bglob * T + (dir ? T - 1 : 0)) * 3 * H + u0 + gu;
        gir = gp[0];
        giz = gp[H];
        gin = gp[2 * H];
    }
    for (int step = 0; step < T; step++) {
        const int t = dir ? (T - 1 - step) : step;
        float nir = 0.f, niz = 0.f, nin = 0.f;
        if (active && step + 1 < T) {
            const int tn = dir ? (T - 2 - step) : step + 1;
            const float* gp = gi + ((size_t)bglob * T + tn) * 3 * H + u0 + gu;
            nir = gp[0];
            niz = gp[H];
            nin = gp[2 * H];
        }
#pragma unroll
        for (int nb = 0; nb < NB; nb++) part[seg][nb][row] = dot64_ffma2<WR>(w, ws4 + tid, NT, &h_s[cur][nb][seg * SEG]);
        __syncthreads();
        const int nxt = CS > 1 ? cur ^ 1 : cur;
        if (is_gate) {
            float ghr = bhr, ghz = bhz, ghn = bhn;
#pragma unroll
            for (int s = 0; s < SEGS; s++) {
                ghr += part[s][gb][gu];
                ghz += part[s][gb][HU + gu];
                ghn += part[s][gb][2 * HU + gu];
            }
            const float r = fast_sigmoid(gir + ghr);
            const float zg = fast_sigmoid(giz + ghz);
            const float n = fast_tanh(gin + r * ghn);
            const float hnew = (1.0f - zg) * n + zg * hval;
            if (active) {
                const size_t bt = (size_t)bglob * T + t;
                if (save) {
                    float* gs = gates + bt * 4 * H + u0 + gu;
                    gs[0] = r;
                    gs[H] = zg;
                    gs[2 * H] = n;
                    gs[3 * H] = ghn;
                    hprev[bt * H + u0 + gu] = hval;
                }
                out[bt * 2 * H + dir * H + u0 + gu] = hnew;
            }
            hval = hnew;
            if (CS > 1) {
                cg::cluster_group cl = cg::this_cluster();
#pragma unroll
                for (int rk = 0; rk < CS; rk++) {
                    float* remote = cl.map_shared_rank(&h_s[nxt][gb][u0 + gu], rk);
                    *remote = hnew;
                }
            } else {
                h_s[nxt][gb][u0 + gu] = hnew;
            }
        }
        if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
        cur = nxt;
        gir = nir;
        giz = niz;
        gin = nin;
    }
}

// ----------------------------------------------------------------------------------------------------------------
template <int H, int CS, int NB>
__global__ void __launch_bounds__(GruCfg<H, CS>::NT_B, 1)
gru_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ whh0, const float* __restrict__ whh1,
               const float* __restrict__ gates0, const float* __restrict__ gates1, const float* __restrict__ hprev0,
               const float* __restrict__ hprev1, float* __restrict__ dgi0, float* __restrict__ dgi1,
               float* __restrict__ dghn0, float* __restrict__ dghn1, float* __restrict__ gbih0,
               float* __restrict__ gbih1, float* __restrict__ gbhh0, float* __restrict__ gbhh1, int B, int T) {
    using Cfg = GruCfg<H, CS>;
    constexpr int HU = Cfg::HU, JSEGS = Cfg::JSEGS, NT = Cfg::NT_B;
    constexpr int WR = WSplit<NT>::WR, WS4 = WSplit<NT>::WS4;
    extern __shared__ __align__(16) float gru_smem[];
    float4* ws4 = reinterpret_cast<float4*>(gru_smem);                                  // [WS4][NT] float4
    // recurrent pre-activation grads (r, z, hn) of all units, double-buffered: [2][NB][3H]
    float (*dgh_s)[NB][3 * H] = reinterpret_cast<float (*)[NB][3 * H]>(gru_smem + WS4 * NT * 4);
    float (*part)[NB][HU] = reinterpret_cast<float (*)[NB][HU]>(gru_smem + WS4 * NT * 4 + 2 * NB * 3 * H);  // [JSEGS][NB][HU]
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    int crank = 0;
    if (CS > 1) crank = (int)cg::this_cluster().block_rank();
    const int b0 = (blockIdx.x / CS) * NB;
    const int u0 = crank * HU;
    const float* whh = dir ? whh1 : whh0;
    const float* gates = dir ? gates1 : gates0;
    const float* hprev = dir ? hprev1 : hprev0;
    float* dgi = dir ? dgi1 : dgi0;
    float* dghn = dir ? dghn1 : dghn0;
    float* gbih = dir ? gbih1 : gbih0;
    float* gbhh = dir ? gbhh1 : gbhh0;
    float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_hn = 0.f;      // bias gradients = sums over time of the gate grads

    const int jseg = tid / HU, col = tid - jseg * HU;        // column u0+col of W_hh, rows jseg*64 .. +64
    float2 w[WR / 2];
#pragma unroll
    for (int k = 0; k < WR / 2; k++)
        w[k] = make_float2(whh[(size_t)(jseg * SEG + 2 * k) * H + u0 + col],
                           whh[(size_t)(jseg * SEG + 2 * k + 1) * H + u0 + col]);
#pragma unroll
    for (int c = 0; c < WS4; c++) {
        const size_t r0 = (size_t)(jseg * SEG + WR + 4 * c) * H + u0 + col;
        ws4[c * NT + tid] = make_float4(whh[r0], whh[r0 + H], whh[r0 + 2 * H], whh[r0 + 3 * H]);
    }

    for (int i = tid; i < 2 * NB * 3 * H; i += NT) (&dgh_s[0][0][0])[i] = 0.f;
    const bool is_gate = tid < NB * HU;
    const int gb = tid / HU, gu = tid - gb * HU;
    const int bglob = b0 + gb;
    const bool active = is_gate && (bglob < B);
    float dh = 0.f;
    if (CS > 1) cg::this_cluster().sync(); else __syncthreads();

    int cur = 0;
    // saved activations of the next processed step are prefetched one step ahead
    float p_go = 0.f, p_r = 0.f, p_z = 0.f, p_n = 0.f, p_ghn = 0.f, p_hp = 0.f;
    if (active) {
        const size_t bt = (size_t)bglob * T + (dir ? 0 : T - 1);
        p_go = gout[bt * 2 * H + dir * H + u0 + gu];
        const float* gs = gates + bt * 4 * H + u0 + gu;
        p_r = gs[0]; p_z = gs[H]; p_n = gs[2 * H]; p_ghn = gs[3 * H];
        p_hp = hprev[bt * H + u0 + gu];
    }
    for (int step = T - 1; step >= 0; step--) {
        const int t = dir ? (T - 1 - step) : step;
        float n_go = 0.f, n_r = 0.f, n_z = 0.f, n_n = 0.f, n_ghn = 0.f, n_hp = 0.f;
        if (active && step > 0) {
            const int tn = dir ? (T - step) : step - 1;
            const size_t btn = (size_t)bglob * T + tn;
            n_go = gout[btn * 2 * H + dir * H + u0 + gu];
            const float* gs = gates + btn * 4 * H + u0 + gu;
            n_r = gs[0]; n_z = gs[H]; n_n = gs[2 * H]; n_ghn = gs[3 * H];
            n_hp = hprev[btn * H + u0 + gu];
        }
        float dh_direct = 0.f, dr_pre = 0.f, dz_pre = 0.f, dn_pre = 0.f, dhn = 0.f;
        if (active) {
            const size_t bt = (size_t)bglob * T + t;
            const float g = p_go + dh;
            const float r = p_r, zg = p_z, n = p_n, ghn = p_ghn;
            const float hp = p_hp;
            const float dn = g * (1.0f - zg);
            const float dz = g * (hp - n);
            dh_direct = g * zg;
            dn_pre = dn * (1.0f - n * n);
            dz_pre = dz * zg * (1.0f - zg);
            dr_pre = dn_pre * ghn * r * (1.0f - r);
            dhn = dn_pre * r;
            float* dp = dgi + bt * 3 * H + u0 + gu;
            dp[0] = dr_pre;
            dp[H] = dz_pre;
            dp[2 * H] = dn_pre;
            dghn[bt * H + u0 + gu] = dhn;
            sb_r += dr_pre;
            sb_z += dz_pre;
            sb_n += dn_pre;
            sb_hn += dhn;
        }
        if (is_gate) {
            if (CS > 1) {
                cg::cluster_group cl = cg::this_cluster();
#pragma unroll
                for (int rk = 0; rk < CS; rk++) {
                    float* base = cl.map_shared_rank(&dgh_s[cur][gb][0], rk);
                    base[u0 + gu] = dr_pre;
                    base[H + u0 + gu] = dz_pre;
                    base[2 * H + u0 + gu] = dhn;
                }
            } else {
                dgh_s[cur][gb][u0 + gu] = dr_pre;
                dgh_s[cur][gb][H + u0 + gu] = dz_pre;
                dgh_s[cur][gb][2 * H + u0 + gu] = dhn;
            }
        }
        if (CS > 1) cg::this_cluster().sync(); else __syncthreads();
        // dh_prev[u] += sum_j W_hh[j][u] * dgh[j]
#pragma unroll
        for (int nb = 0; nb < NB; nb++) part[jseg][nb][col] = dot64_ffma2<WR>(w, ws4 + tid, NT, &dgh_s[cur][nb][jseg * SEG]);
        __syncthreads();
        if (is_gate) {
            float s = dh_direct;
#pragma unroll
            for (int js = 0; js < JSEGS; js++) s += part[js][gb][gu];
            dh = s;
        }
        p_go = n_go; p_r = n_r; p_z = n_z; p_n = n_n; p_ghn = n_ghn; p_hp = n_hp;
        if (CS > 1) cur ^= 1;      // next step's remote writes must not race with slower CTAs still reading
    }
    if (active && gbih != nullptr) {
        // b_ih and b_hh share the r and z gradients; the n gate differs (d n_pre vs d(hn) = d n_pre * r)
        atomicAdd(&gbih[u0 + gu], sb_r);
        atomicAdd(&gbih[H + u0 + gu], sb_z);
        atomicAdd(&gbih[2 * H + u0 + gu], sb_n);
        atomicAdd(&gbhh[u0 + gu], sb_r);
        atomicAdd(&gbhh[H + u0 + gu], sb_z);
        atomicAdd(&gbhh[2 * H + u0 + gu], sb_hn);
    }
}

template <int H, int CS, int NB>
int run_fwd(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
            float* const gates[2], float* const hprev[2], int B, int T, int save, cudaStream_t s) {
    using Cfg = GruCfg<H, CS>;
    static_assert(GruNbOk<H, CS, NB>::ok, "");
    auto kern = gru_fwd_kernel<H, CS, NB>;
    dim3 grid(cdiv(B, NB) * CS, 2);
    const size_t smem = (size_t)(WSplit<Cfg::NT_F>::WS4 * Cfg::NT_F * 4 + 2 * NB * H + Cfg::SEGS * NB * Cfg::R) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        int rc = opt_in_smem(kern, smem);
        if (rc) return rc;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(Cfg::NT_F);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SEDK_CUDA(cudaLaunchKernelEx(&cfg, kern, gi[0], gi[1], w_hh[0], w_hh[1], b_hh[0], b_hh[1], out, gates[0], gates[1],
                                 hprev[0], hprev[1], B, T, save));
    count_launch();
    return SEDK_OK;
}

template <int H, int CS, int NB>
int run_bwd(const float* gout, const float* const w_hh[2], const float* const gates[2], const float* const hprev[2],
            float* const dgi[2], float* const dghn[2], float* const gb_ih[2], float* const gb_hh[2], int B, int T,
            cudaStream_t s) {
    using Cfg = GruCfg<H, CS>;
    auto kern = gru_bwd_kernel<H, CS, NB>;
    dim3 grid(cdiv(B, NB) * CS, 2);
    const size_t smem = (size_t)(WSplit<Cfg::NT_B>::WS4 * Cfg::NT_B * 4 + 2 * NB * 3 * H + Cfg::JSEGS * NB * Cfg::HU) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        int rc = opt_in_smem(kern, smem);
        if (rc) return rc;
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(Cfg::NT_B);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int d = 0; d < 2; d++) {
        SEDK_CUDA(cudaMemsetAsync(gb_ih[d], 0, (size_t)3 * H * sizeof(float), s));
        SEDK_CUDA(cudaMemsetAsync(gb_hh[d], 0, (size_t)3 * H * sizeof(float), s));
    }
    SEDK_CUDA(cudaLaunchKernelEx(&cfg, kern, gout, w_hh[0], w_hh[1], gates[0], gates[1], hprev[0], hprev[1], dgi[0],
                                 dgi[1], dghn[0], dghn[1], gb_ih[0], gb_ih[1], gb_hh[0], gb_hh[1], B, T));
    count_launch();
    return SEDK_OK;
}

// H = 128 can also run as a 2-CTA cluster (all weights in registers, half the issue work per SM, one DSMEM exchange
// + cluster barrier per step); selected with sedk_set_gru_cluster(2) / env SEDK_GRU_CLUSTER=2 for A/B measurements
int gru_cluster();
}  // namespace
int g_gru_cluster = -1;
namespace {
int gru_cluster() {
    if (g_gru_cluster < 0) {
        const char* e = getenv("SEDK_GRU_CLUSTER");
        g_gru_cluster = (e != nullptr && e[0] == '2') ? 2 : 1;
    }
    return g_gru_cluster;
}

// batch rows per CTA: keep every (row, direction) pair on its own SM while they fit, then double up
inline int pick_nb(int B, int CS) {
    const int sms = num_sms();
    int nb = 1;
    while (nb < 4 && cdiv(B, nb) * 2 * CS > sms) nb *= 2;
    return nb;
}

}  // namespace

int launch_gru_seq_fwd(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
                       float* const gates[2], float* const hprev[2], int B, int T, int H, int save, cudaStream_t s) {
    SEDK_PROF("gru_seq_fwd", s);
    if (H == 128 && gru_cluster() == 2 && pick_nb(B, 2) == 1)
        return run_fwd<128, 2, 1>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
    if (H == 128) {
        switch (pick_nb(B, 1)) {
            case 1: return run_fwd<128, 1, 1>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
            case 2: return run_fwd<128, 1, 2>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
            default: return run_fwd<128, 1, 4>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
        }
    }
    if (H == 64) return run_fwd<64, 1, 2>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
    if (H == 192) {
        switch (pick_nb(B, 3)) {
            case 1: return run_fwd<192, 3, 1>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
            case 2: return run_fwd<192, 3, 2>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
            default: return run_fwd<192, 3, 4>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
        }
    }
    SEDK_UNSUPPORTED("GRU hidden size %d has no sm_100a instantiation (supported: 64, 128, 192)", H);
}

int launch_gru_seq_bwd(const float* gout, const float* const w_hh[2], const float* const gates[2],
                       const float* const hprev[2], float* const dgi[2], float* const dghn[2], float* const gb_ih[2],
                       float* const gb_hh[2], int B, int T, int H, cudaStream_t s) {
    SEDK_PROF("gru_seq_bwd", s);
    if (H == 128 && gru_cluster() == 2 && pick_nb(B, 2) == 1)
        return run_bwd<128, 2, 1>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
    if (H == 128) {
        switch (pick_nb(B, 1)) {
            case 1: return run_bwd<128, 1, 1>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
            case 2: return run_bwd<128, 1, 2>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
            default: return run_bwd<128, 1, 4>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
        }
    }
    if (H == 64) return run_bwd<64, 1, 2>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
    if (H == 192) {
        switch (pick_nb(B, 3)) {
            case 1: return run_bwd<192, 3, 1>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
            case 2: return run_bwd<192, 3, 2>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
            default: return run_bwd<192, 3, 4>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, s);
        }
    }
    SEDK_UNSUPPORTED("GRU hidden size %d has no sm_100a instantiation (supported: 64, 128, 192)", H);
}

}  // namespace sedk

extern "C" int sedk_set_gru_cluster(int cs) {
    sedk::g_gru_cluster = (cs == 2) ? 2 : 1;
    return SEDK_OK;
}
