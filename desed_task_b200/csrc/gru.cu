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

// ================================================================================================================
// Second-generation recurrence for H = 128 ("v2"), one batch row per CTA.  ncu on the first-generation kernel: 1460
// shared-memory wavefronts per time step (h broadcast + smem-resident weights + partial sums) and two barriers, ~2100
// cycles per step.  v2:
//   forward : 512 threads = 128 hidden units x 4 k-lanes.  A quad owns the r, z, n rows of one unit; lane kl holds the
//             3 x 32 weights of k in [32 kl, 32 kl + 32) (80 in registers, 16 in shared memory), reads its 32 h values as
//             8 LDS.128 (h is stored with 4 floats of padding per 32 so the four lanes of a quad hit different banks),
//             reduces the three partial sums over the quad with 6 shuffles and finishes the cell redundantly in all four
//             lanes.  The new h goes to the other half of a double buffer: ONE __syncthreads per step.
//   backward: 512 threads = 64 unit pairs x 8 j-lanes.  An octet owns columns (2o, 2o+1) of W_hh; lane l8 holds the
//             2 x 48 weights of rows j in [48 l8, 48 l8 + 48) and reads dgh[j] as 12 LDS.128 (padding 4 per 48);
//             a 3-shuffle exchange/reduce leaves d h_prev of unit 2o in lanes 0-3 and of unit 2o+1 in lanes 4-7, where
//             the gate gradients of the next (earlier) step are computed - no shared-memory partials, one barrier per step.
// Per-step integer work is kept to pointer increments: every global array is walked with a per-lane pointer and a
// per-lane stride, and each lane of a quad owns two of the six stores of its unit.
constexpr int V2_H = 128;
constexpr int V2_NT = 512;
constexpr int V2_WS4 = 4;            // float4 chunks of weights per thread kept in shared memory (16 of 96 weights)
constexpr int V2_HPAD = 36 * 4;      // padded h buffer: k -> k + 4 (k / 32)
constexpr int V2_DPAD = 52 * 8;      // padded dgh buffer: j -> j + 4 (j / 48)

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
// MUFU-only gate nonlinearities without the range-check code of __expf / __fdividef (inf / 0 propagate to the right limits)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lean_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float lean_tanh(float x) {
    return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * x)), 1.0f);
}

__global__ void __launch_bounds__(V2_NT, 1)
gru_fwd_v2_kernel(const float* __restrict__ gi0, const float* __restrict__ gi1, const float* __restrict__ whh0,
                  const float* __restrict__ whh1, const float* __restrict__ bhh0, const float* __restrict__ bhh1,
                  float* __restrict__ out, float* __restrict__ gates0, float* __restrict__ gates1,
                  float* __restrict__ hprev0, float* __restrict__ hprev1, int T, int save) {
    constexpr int H = V2_H, NT = V2_NT;
    extern __shared__ __align__(16) float gru_smem[];
    float4* ws4 = reinterpret_cast<float4*>(gru_smem);            // [V2_WS4][NT]
    float* h_s = gru_smem + V2_WS4 * NT * 4;                      // [2][V2_HPAD]
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const int b = blockIdx.x;
    const int u = tid >> 2, kl = tid & 3;
    const float* whh = dir ? whh1 : whh0;
    const float* bhh = dir ? bhh1 : bhh0;

    float2 wr[16], wz[16], wn[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int k = 32 * kl + 4 * c;
        const float4 a = *reinterpret_cast<const float4*>(whh + (size_t)u * H + k);
        const float4 bz = *reinterpret_cast<const float4*>(whh + (size_t)(H + u) * H + k);
        const float4 n = *reinterpret_cast<const float4*>(whh + (size_t)(2 * H + u) * H + k);
        wr[2 * c] = lo2(a); wr[2 * c + 1] = hi2(a);
        wz[2 * c] = lo2(bz); wz[2 * c + 1] = hi2(bz);
        if (c < 4) { wn[2 * c] = lo2(n); wn[2 * c + 1] = hi2(n); }
        else ws4[(c - 4) * NT + tid] = n;
    }
    for (int i = tid; i < 2 * V2_HPAD; i += NT) h_s[i] = 0.f;
    const float bhr = bhh[u], bhz = bhh[H + u], bhn = bhh[2 * H + u];

    // per-lane walking pointers (time stride +-1 step)
    const int t0 = dir ? T - 1 : 0;
    const ptrdiff_t ts = dir ? -1 : 1;
    const size_t bt0 = (size_t)b * T + t0;
    const float* gp = (dir ? gi1 : gi0) + bt0 * 3 * H + u;
    const ptrdiff_t gstep = ts * 3 * H;
    // lane 0: out <- h_new, hprev <- h_old;  lane 1: gates r, z;  lane 2: gates n, hn;  lane 3: nothing
    float* gbase = (dir ? gates1 : gates0) + bt0 * 4 * H + u;
    float* pa = kl == 0 ? out + bt0 * 2 * H + dir * H + u : (kl == 1 ? gbase : gbase + 2 * H);
    float* pb = kl == 0 ? (dir ? hprev1 : hprev0) + bt0 * H + u : (kl == 1 ? gbase + H : gbase + 3 * H);
    const ptrdiff_t sa = ts * (kl == 0 ? 2 * H : 4 * H), sb = ts * (kl == 0 ? H : 4 * H);
    const bool do_a = kl == 0 || (save != 0 && kl < 3);
    const bool do_b = save != 0 && kl < 3;
    const float* hrd = h_s + 36 * kl;                             // this lane's 32 h values (padded layout), buffer 0
    float* hwr = h_s + V2_HPAD + u + 4 * (u >> 5);                // where lane 0 of the quad publishes h_new, buffer 1
    float hval = 0.f;
    float gir = gp[0] + bhr, giz = gp[H] + bhz, gin = gp[2 * H];
    __syncthreads();

    for (int step = 0; step < T; step++) {
        float nir = 0.f, niz = 0.f, nin = 0.f;
        if (step + 1 < T) {
            gp += gstep;
            nir = gp[0]; niz = gp[H]; nin = gp[2 * H];
        }
        float2 ar0 = make_float2(0.f, 0.f), ar1 = ar0, az0 = ar0, az1 = ar0, an0 = ar0, an1 = ar0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const float4 h4 = *reinterpret_cast<const float4*>(hrd + 4 * c);
            const float2 hl = lo2(h4), hh = hi2(h4);
            ar0 = __ffma2_rn(wr[2 * c], hl, ar0);
            ar1 = __ffma2_rn(wr[2 * c + 1], hh, ar1);
            az0 = __ffma2_rn(wz[2 * c], hl, az0);
            az1 = __ffma2_rn(wz[2 * c + 1], hh, az1);
            if (c < 4) {
                an0 = __ffma2_rn(wn[2 * c], hl, an0);
                an1 = __ffma2_rn(wn[2 * c + 1], hh, an1);
            } else {
                const float4 w4 = ws4[(c - 4) * NT + tid];
                an0 = __ffma2_rn(lo2(w4), hl, an0);
                an1 = __ffma2_rn(hi2(w4), hh, an1);
            }
        }
        float sr = (ar0.x + ar0.y) + (ar1.x + ar1.y);
        float sz = (az0.x + az0.y) + (az1.x + az1.y);
        float sn = (an0.x + an0.y) + (an1.x + an1.y);
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            sr += __shfl_xor_sync(0xffffffffu, sr, o);
            sz += __shfl_xor_sync(0xffffffffu, sz, o);
            sn += __shfl_xor_sync(0xffffffffu, sn, o);
        }
        const float ghn = sn + bhn;
        const float r = lean_sigmoid(gir + sr);
        const float zg = lean_sigmoid(giz + sz);
        const float n = lean_tanh(fmaf(r, ghn, gin));
        const float hnew = fmaf(zg, hval - n, n);                 // (1 - z) n + z h
        if (kl == 0) *hwr = hnew;
        if (do_a) *pa = kl == 0 ? hnew : (kl == 1 ? r : n);
        if (do_b) *pb = kl == 0 ? hval : (kl == 1 ? zg : ghn);
        pa += sa;
        pb += sb;
        hval = hnew;
        gir = nir + bhr; giz = niz + bhz; gin = nin;
        // swap the double buffer: readers move to the buffer just written
        const ptrdiff_t flip = (step & 1) ? -V2_HPAD : V2_HPAD;
        hrd += flip;
        hwr -= flip;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(V2_NT, 1)
gru_bwd_v2_kernel(const float* __restrict__ gout, const float* __restrict__ whh0, const float* __restrict__ whh1,
                  const float* __restrict__ gates0, const float* __restrict__ gates1, const float* __restrict__ hprev0,
                  const float* __restrict__ hprev1, float* __restrict__ dgi0, float* __restrict__ dgi1,
                  float* __restrict__ dghn0, float* __restrict__ dghn1, float* __restrict__ gbih0,
                  float* __restrict__ gbih1, float* __restrict__ gbhh0, float* __restrict__ gbhh1, int T) {
    constexpr int H = V2_H, NT = V2_NT;
    extern __shared__ __align__(16) float gru_smem[];
    float4* ws4 = reinterpret_cast<float4*>(gru_smem);            // [V2_WS4][NT]
    float* dgh_s = gru_smem + V2_WS4 * NT * 4;                    // [2][V2_DPAD]: d(r_pre), d(z_pre), d(hn), padded
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const int b = blockIdx.x;
    const int o = tid >> 3, l8 = tid & 7;
    const int ua = 2 * o;                       // columns ua, ua + 1 of W_hh
    const int um = ua + (l8 >> 2);              // the unit whose cell gradient this lane computes
    const int sub = l8 & 3;
    const float* whh = dir ? whh1 : whh0;
    float* gbih = dir ? gbih1 : gbih0;
    float* gbhh = dir ? gbhh1 : gbhh0;

    float2 wa[24], wb[16];
#pragma unroll
    for (int c = 0; c < 12; c++) {
        const float* wp = whh + (size_t)(48 * l8 + 4 * c) * H + ua;
        const float2 r0 = *reinterpret_cast<const float2*>(wp);
        const float2 r1 = *reinterpret_cast<const float2*>(wp + H);
        const float2 r2 = *reinterpret_cast<const float2*>(wp + 2 * H);
        const float2 r3 = *reinterpret_cast<const float2*>(wp + 3 * H);
        wa[2 * c] = make_float2(r0.x, r1.x);
        wa[2 * c + 1] = make_float2(r2.x, r3.x);
        if (c < 8) {
            wb[2 * c] = make_float2(r0.y, r1.y);
            wb[2 * c + 1] = make_float2(r2.y, r3.y);
        } else {
            ws4[(c - 8) * NT + tid] = make_float4(r0.y, r1.y, r2.y, r3.y);
        }
    }
    for (int i = tid; i < 2 * V2_DPAD; i += NT) dgh_s[i] = 0.f;
    float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_hn = 0.f;

    // processing order: time index t = dir ? T-1-step : step for step = T-1 .. 0
    const int t0 = dir ? 0 : T - 1;
    const ptrdiff_t ts = dir ? 1 : -1;
    const size_t bt0 = (size_t)b * T + t0;
    const float* gop = gout + bt0 * 2 * H + dir * H + um;
    const float* gsp = (dir ? gates1 : gates0) + bt0 * 4 * H + um;
    const float* hpp = (dir ? hprev1 : hprev0) + bt0 * H + um;
    // lane sub 0: dgi r, z;  sub 1: dgi n, dghn;  sub 2: the three shared-memory values;  sub 3: nothing
    float* dgb = (dir ? dgi1 : dgi0) + bt0 * 3 * H + um;
    float* pa = sub == 0 ? dgb : dgb + 2 * H;
    float* pb = sub == 0 ? dgb + H : (dir ? dghn1 : dghn0) + bt0 * H + um;
    const ptrdiff_t sa = ts * 3 * H, sbs = ts * (sub == 0 ? 3 * H : H);
    const bool do_g = sub < 2;
    // padded positions of (r, z, hn) of unit um in the dgh buffer, and of this lane's 48-row slice
    const int jr = um, jz = H + um, jn = 2 * H + um;
    float* dwr = dgh_s + jr + 4 * (jr / 48);
    const int offz = (jz + 4 * (jz / 48)) - (jr + 4 * (jr / 48)), offn = (jn + 4 * (jn / 48)) - (jr + 4 * (jr / 48));
    const float* drd = dgh_s + 52 * l8;
    float dh = 0.f;
    float p_go = gop[0], p_r = gsp[0], p_z = gsp[H], p_n = gsp[2 * H], p_ghn = gsp[3 * H], p_hp = hpp[0];
    __syncthreads();

    for (int step = T - 1; step >= 0; step--) {
        // saved activations of the next processed step, one step ahead of their use
        float n_go = 0.f, n_r = 0.f, n_z = 0.f, n_n = 0.f, n_ghn = 0.f, n_hp = 0.f;
        if (step > 0) {
            gop += ts * 2 * H;
            gsp += ts * 4 * H;
            hpp += ts * H;
            n_go = gop[0];
            n_r = gsp[0]; n_z = gsp[H]; n_n = gsp[2 * H]; n_ghn = gsp[3 * H];
            n_hp = hpp[0];
        }
        const float g = p_go + dh;
        const float dn = g * (1.0f - p_z);
        const float dz = g * (p_hp - p_n);
        const float dh_direct = g * p_z;
        const float dn_pre = dn * (1.0f - p_n * p_n);
        const float dz_pre = dz * p_z * (1.0f - p_z);
        const float dr_pre = dn_pre * p_ghn * p_r * (1.0f - p_r);
        const float dhn = dn_pre * p_r;
        if (do_g) {
            *pa = sub == 0 ? dr_pre : dn_pre;
            *pb = sub == 0 ? dz_pre : dhn;
        } else if (sub == 2) {
            dwr[0] = dr_pre;
            dwr[offz] = dz_pre;
            dwr[offn] = dhn;
        }
        pa += sa;
        pb += sbs;
        sb_r += dr_pre; sb_z += dz_pre; sb_n += dn_pre; sb_hn += dhn;
        p_go = n_go; p_r = n_r; p_z = n_z; p_n = n_n; p_ghn = n_ghn; p_hp = n_hp;
        __syncthreads();
        // d h_prev[u] = dh_direct[u] + sum_j W_hh[j][u] * dgh[j]
        float2 aa0 = make_float2(0.f, 0.f), aa1 = aa0, ab0 = aa0, ab1 = aa0;
#pragma unroll
        for (int c = 0; c < 12; c++) {
            const float4 d4 = *reinterpret_cast<const float4*>(drd + 4 * c);
            const float2 dl = lo2(d4), dhh = hi2(d4);
            aa0 = __ffma2_rn(wa[2 * c], dl, aa0);
            aa1 = __ffma2_rn(wa[2 * c + 1], dhh, aa1);
            if (c < 8) {
                ab0 = __ffma2_rn(wb[2 * c], dl, ab0);
                ab1 = __ffma2_rn(wb[2 * c + 1], dhh, ab1);
            } else {
                const float4 w4 = ws4[(c - 8) * NT + tid];
                ab0 = __ffma2_rn(lo2(w4), dl, ab0);
                ab1 = __ffma2_rn(hi2(w4), dhh, ab1);
            }
        }
        const float sa_ = (aa0.x + aa0.y) + (aa1.x + aa1.y);
        const float sb_ = (ab0.x + ab0.y) + (ab1.x + ab1.y);
        const bool upper = (l8 & 4) != 0;
        float keep = upper ? sb_ : sa_;
        keep += __shfl_xor_sync(0xffffffffu, upper ? sa_ : sb_, 4);
        keep += __shfl_xor_sync(0xffffffffu, keep, 2);
        keep += __shfl_xor_sync(0xffffffffu, keep, 1);
        dh = dh_direct + keep;
        // the next step's gate writes go to the other buffer: no second barrier needed
        const ptrdiff_t flip = ((T - 1 - step) & 1) ? -V2_DPAD : V2_DPAD;
        drd += flip;
        dwr += flip;
    }
    if (sub == 0 && gbih != nullptr) {
        // b_ih and b_hh share the r and z gradients; the n gate differs (d n_pre vs d(hn) = d n_pre * r)
        atomicAdd(&gbih[um], sb_r);
        atomicAdd(&gbih[H + um], sb_z);
        atomicAdd(&gbih[2 * H + um], sb_n);
        atomicAdd(&gbhh[um], sb_r);
        atomicAdd(&gbhh[H + um], sb_z);
        atomicAdd(&gbhh[2 * H + um], sb_hn);
    }
}

int run_fwd_v2(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
               float* const gates[2], float* const hprev[2], int B, int T, int save, cudaStream_t s) {
    auto kern = gru_fwd_v2_kernel;
    const size_t smem = (size_t)(V2_WS4 * V2_NT * 4 + 2 * V2_HPAD) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        int rc = opt_in_smem(kern, smem);
        if (rc) return rc;
        configured = true;
    }
    kern<<<dim3(B, 2), V2_NT, smem, s>>>(gi[0], gi[1], w_hh[0], w_hh[1], b_hh[0], b_hh[1], out, gates[0], gates[1],
                                         hprev[0], hprev[1], T, save);
    SEDK_LAUNCH_CHECK("gru_fwd_v2_kernel");
    return SEDK_OK;
}

int run_bwd_v2(const float* gout, const float* const w_hh[2], const float* const gates[2], const float* const hprev[2],
               float* const dgi[2], float* const dghn[2], float* const gb_ih[2], float* const gb_hh[2], int B, int T,
               int zeroed, cudaStream_t s) {
    auto kern = gru_bwd_v2_kernel;
    const size_t smem = (size_t)(V2_WS4 * V2_NT * 4 + 2 * V2_DPAD) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        int rc = opt_in_smem(kern, smem);
        if (rc) return rc;
        configured = true;
    }
    for (int d = 0; d < 2 && !zeroed; d++) {
        SEDK_CUDA(cudaMemsetAsync(gb_ih[d], 0, (size_t)3 * V2_H * sizeof(float), s));
        SEDK_CUDA(cudaMemsetAsync(gb_hh[d], 0, (size_t)3 * V2_H * sizeof(float), s));
    }
    kern<<<dim3(B, 2), V2_NT, smem, s>>>(gout, w_hh[0], w_hh[1], gates[0], gates[1], hprev[0], hprev[1], dgi[0], dgi[1],
                                         dghn[0], dghn[1], gb_ih[0], gb_ih[1], gb_hh[0], gb_hh[1], T);
    SEDK_LAUNCH_CHECK("gru_bwd_v2_kernel");
    return SEDK_OK;
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
            int zeroed, cudaStream_t s) {
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
    for (int d = 0; d < 2 && !zeroed; d++) {
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
    // v3 (gru3.cu): octet-per-unit-group layout, every weight in registers, 4 LDS.128 of h per thread and step
    if (H == 128 && get_option("gru_v3", 1) && pick_nb(B, 1) == 1)
        return launch_gru_fwd_v3(gi, w_hh, b_hh, out, gates, hprev, B, T, save, get_option("gru_v3", 1), s);
    // v2 holds 80 weights + the loop state in exactly 128 registers at one batch row per CTA (two rows spill)
    if (H == 128 && get_option("gru_v2", 1) && pick_nb(B, 1) == 1)
        return run_fwd_v2(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
    if (H == 128) {
        switch (pick_nb(B, 1)) {
            case 1: return run_fwd<128, 1, 1>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
            case 2: return run_fwd<128, 1, 2>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
            default: return run_fwd<128, 1, 4>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
        }
    }
    if (H == 64) return run_fwd<64, 1, 2>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
    if (H == 192 && get_option("gru_v3", 1)) return launch_gru_fwd_c3(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
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
                       float* const gb_hh[2], int B, int T, int H, int zeroed, cudaStream_t s) {
    SEDK_PROF("gru_seq_bwd", s);
    if (H == 128 && gru_cluster() == 2 && pick_nb(B, 2) == 1)
        return run_bwd<128, 2, 1>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
    if (H == 128 && get_option("gru_v3", 1) && pick_nb(B, 1) == 1)
        return launch_gru_bwd_v3(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, get_option("gru_v3", 1), s);
    if (H == 128 && get_option("gru_v2", 1) && pick_nb(B, 1) == 1)
        return run_bwd_v2(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
    if (H == 128) {
        switch (pick_nb(B, 1)) {
            case 1: return run_bwd<128, 1, 1>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
            case 2: return run_bwd<128, 1, 2>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
            default: return run_bwd<128, 1, 4>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
        }
    }
    if (H == 64) return run_bwd<64, 1, 2>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
    if (H == 192 && get_option("gru_v3", 1))
        return launch_gru_bwd_c3(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
    if (H == 192) {
        switch (pick_nb(B, 3)) {
            case 1: return run_bwd<192, 3, 1>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
            case 2: return run_bwd<192, 3, 2>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
            default: return run_bwd<192, 3, 4>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
        }
    }
    SEDK_UNSUPPORTED("GRU hidden size %d has no sm_100a instantiation (supported: 64, 128, 192)", H);
}

}  // namespace sedk

extern "C" int sedk_set_gru_cluster(int cs) {
    sedk::g_gru_cluster = (cs == 2) ? 2 : 1;
    return SEDK_OK;
}
