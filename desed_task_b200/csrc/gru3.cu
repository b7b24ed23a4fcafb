// Third-generation persistent GRU recurrence for H = 128 (nn.GRU semantics, gate order r,z,n; desed_task/nnet/RNN.py:19-30).
//
// Same decomposition as gru.cu: the input-side GEMMs are hoisted, one CTA owns one (batch row, direction) and keeps W_hh
// resident for all T steps, exact fp32 accumulation (packed FFMA2).  What changed is WHERE the per-step time went.
// Measured on B200 (tools/micro/ubench.cu): an LDS.128 costs 4 shared-memory cycles per warp whatever the lanes read (it
// is issued in four quarter-warp phases), SHFL ~4 cycles per SM sub-partition, a block barrier round trip ~100 cycles.
// The second generation re-read h eight times per thread (quad-per-unit layout: 16 warps x 8 LDS.128 = 512 cycles of the
// shared-memory pipe per step, + 256 for the weights that did not fit in registers; backward 768 + 256) on top of the
// 384 cycles the FMA pipe needs for the 49 152 MACs of a step.  Here:
//   forward : an OCTET of lanes owns UPO hidden units (their r, z, n rows); lane l8 holds the k-slice [16 l8, 16 l8 + 16)
//             of those 3 UPO rows and reads its 16 h values with 4 LDS.128 (conflict-free through the 4-per-32 padding).
//             The 3 UPO partial sums are reduced over the octet with a transposing butterfly (each level halves the number
//             of live values), the gate math runs once per unit pair of lanes.  UPO = 4: 8 warps, all 192 weights of a
//             thread in registers, no weight traffic at all; h costs 8 warps x 4 x 4 = 128 shared-memory cycles per step.
//   backward: the 384 gate-gradient rows are split over ALL 32 lanes of a warp (12 each: three conflict-free LDS.128), a
//             warp owns UPW columns of W_hh; the UPW column sums are reduced with the same transposing butterfly.
//             UPW = 16: 8 warps, 192 weights per thread in registers; dgh costs 8 x 3 x 4 = 96 cycles per step.
// One __syncthreads per step (double-buffered h / dgh), every global array is walked with per-lane pointers.
#include "kernels.h"

namespace sedk {
namespace {

constexpr int H3 = 128;
constexpr int HPAD3 = 144;           // padded h buffer: k -> k + 4 (k / 32)

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
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

// one level of the transposing butterfly: N live values per lane -> N / 2; lanes with `up` keep the upper half
template <int N>
__device__ __forceinline__ void fold(float* v, bool up, int xor_mask) {
#pragma unroll
    for (int j = 0; j < N / 2; j++) {
        const float send = up ? v[j] : v[j + N / 2];
        const float keep = up ? v[j + N / 2] : v[j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, xor_mask);
    }
}

// ----------------------------------------------------------------------------------------------------------------
template <int UPO>
__global__ void __launch_bounds__(128 / UPO * 8, 1)
gru_fwd_v3_kernel(const float* __restrict__ gi0, const float* __restrict__ gi1, const float* __restrict__ whh0,
                  const float* __restrict__ whh1, const float* __restrict__ bhh0, const float* __restrict__ bhh1,
                  float* __restrict__ out, float* __restrict__ gates0, float* __restrict__ gates1,
                  float* __restrict__ hprev0, float* __restrict__ hprev1, int T, int save) {
    pdl_enter();
    constexpr int H = H3, NT = 128 / UPO * 8, R = 3 * UPO;
    constexpr int LPU = 8 / UPO;                                  // lanes that end up with the sums of one unit
    __shared__ __align__(16) float h_s[2 * HPAD3];
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const int b = blockIdx.x;
    const int l8 = tid & 7;
    const int ub = (tid >> 3) * UPO;
    const float* whh = dir ? whh1 : whh0;
    const float* bhh = dir ? bhh1 : bhh0;

    // w[i * 3 + g][c] = W_hh[g H + ub + i][16 l8 + 2 c .. + 2]
    float2 w[R][8];
#pragma unroll
    for (int i = 0; i < UPO; i++)
#pragma unroll
        for (int g = 0; g < 3; g++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float4 a = *reinterpret_cast<const float4*>(whh + (size_t)(g * H + ub + i) * H + 16 * l8 + 4 * c);
                w[i * 3 + g][2 * c] = lo2(a);
                w[i * 3 + g][2 * c + 1] = hi2(a);
            }
    for (int i = tid; i < 2 * HPAD3; i += NT) h_s[i] = 0.f;
    // the unit whose cell this lane finishes (after the butterfly) and its rank among the LPU lanes that share it
    const int ul = UPO == 4 ? (l8 >> 1) : (l8 >> 2);
    const int sub = UPO == 4 ? (l8 & 1) : (l8 & 3);
    const int u = ub + ul;
    const float bhr = bhh[u], bhz = bhh[H + u], bhn = bhh[2 * H + u];

    const int t0 = dir ? T - 1 : 0;
    const ptrdiff_t ts = dir ? -1 : 1;
    const size_t bt0 = (size_t)b * T + t0;
    const float* gp = (dir ? gi1 : gi0) + bt0 * 3 * H + u;
    const ptrdiff_t gstep = ts * 3 * H;
    float* const obase = out + bt0 * 2 * H + dir * H + u;
    float* const gbase = (dir ? gates1 : gates0) + bt0 * 4 * H + u;
    float* const hbase = (dir ? hprev1 : hprev0) + bt0 * H + u;
    // store duty of a lane.  LPU = 2: sub 0 -> {out, r, z}, sub 1 -> {hprev, n, hn};  LPU = 4: sub 0 -> {out, hprev},
    // sub 1 -> {r, z}, sub 2 -> {n, hn}, sub 3 -> nothing
    float *pa, *pb, *pc = nullptr;
    ptrdiff_t sa, sb, sc = 0;
    bool do_a, do_b, do_c = false;
    if (LPU == 2) {
        pa = sub == 0 ? obase : hbase;
        sa = ts * (sub == 0 ? 2 * H : H);
        pb = sub == 0 ? gbase : gbase + 2 * H;
        pc = sub == 0 ? gbase + H : gbase + 3 * H;
        sb = sc = ts * 4 * H;
        do_a = sub == 0 || save != 0;
        do_b = do_c = save != 0;
    } else {
        pa = sub == 0 ? obase : (sub == 1 ? gbase : gbase + 2 * H);
        pb = sub == 0 ? hbase : (sub == 1 ? gbase + H : gbase + 3 * H);
        sa = ts * (sub == 0 ? 2 * H : 4 * H);
        sb = ts * (sub == 0 ? H : 4 * H);
        do_a = sub == 0 || (save != 0 && sub < 3);
        do_b = save != 0 && sub < 3;
    }
    const float* hrd = h_s + 16 * l8 + 4 * (l8 >> 1);             // this lane's 16 h values (padded layout), buffer 0
    float* hwr = h_s + HPAD3 + u + 4 * (u >> 5);                  // where sub 0 publishes h_new, buffer 1
    float hval = 0.f;
    float gir = gp[0] + bhr, giz = gp[H] + bhz, gin = gp[2 * H];
    __syncthreads();

    for (int step = 0; step < T; step++) {
        float nir = 0.f, niz = 0.f, nin = 0.f;
        if (step + 1 < T) {
            gp += gstep;
            nir = gp[0]; niz = gp[H]; nin = gp[2 * H];
        }
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const float4 h4 = *reinterpret_cast<const float4*>(hrd + 4 * c);
            const float2 hl = lo2(h4), hh = hi2(h4);
#pragma unroll
            for (int r = 0; r < R; r++) {
                acc[r] = __ffma2_rn(w[r][2 * c], hl, acc[r]);
                acc[r] = __ffma2_rn(w[r][2 * c + 1], hh, acc[r]);
            }
        }
        float v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = acc[r].x + acc[r].y;
        float sr, sz, sn;
        if (UPO == 4) {
            fold<12>(v, (l8 & 4) != 0, 4);
            fold<6>(v, (l8 & 2) != 0, 2);
            sr = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
            sz = v[1] + __shfl_xor_sync(0xffffffffu, v[1], 1);
            sn = v[2] + __shfl_xor_sync(0xffffffffu, v[2], 1);
        } else {
            fold<6>(v, (l8 & 4) != 0, 4);
            sr = v[0]; sz = v[1]; sn = v[2];
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                sr += __shfl_xor_sync(0xffffffffu, sr, o);
                sz += __shfl_xor_sync(0xffffffffu, sz, o);
                sn += __shfl_xor_sync(0xffffffffu, sn, o);
            }
        }
        const float ghn = sn + bhn;
        const float r = lean_sigmoid(gir + sr);
        const float zg = lean_sigmoid(giz + sz);
        const float n = lean_tanh(fmaf(r, ghn, gin));
        const float hnew = fmaf(zg, hval - n, n);                 // (1 - z) n + z h
        if (sub == 0) *hwr = hnew;
        if (LPU == 2) {
            if (do_a) *pa = sub == 0 ? hnew : hval;
            if (do_b) *pb = sub == 0 ? r : n;
            if (do_c) *pc = sub == 0 ? zg : ghn;
            pc += sc;
        } else {
            if (do_a) *pa = sub == 0 ? hnew : (sub == 1 ? r : n);
            if (do_b) *pb = sub == 0 ? hval : (sub == 1 ? zg : ghn);
        }
        pa += sa;
        pb += sb;
        hval = hnew;
        gir = nir + bhr; giz = niz + bhz; gin = nin;
        const ptrdiff_t flip = (step & 1) ? -HPAD3 : HPAD3;       // readers move to the buffer just written
        hrd += flip;
        hwr -= flip;
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------------------------------------------
// d h_prev[u] = dh_direct[u] + sum_j W_hh[j][u] dgh[j],  dgh = [d r_pre (H) | d z_pre (H) | d hn (H)]
template <int UPW>
__global__ void __launch_bounds__(128 / UPW * 32, 1)
gru_bwd_v3_kernel(const float* __restrict__ gout, const float* __restrict__ whh0, const float* __restrict__ whh1,
                  const float* __restrict__ gates0, const float* __restrict__ gates1, const float* __restrict__ hprev0,
                  const float* __restrict__ hprev1, float* __restrict__ dgi0, float* __restrict__ dgi1,
                  float* __restrict__ dghn0, float* __restrict__ dghn1, float* __restrict__ gbih0,
                  float* __restrict__ gbih1, float* __restrict__ gbhh0, float* __restrict__ gbhh1, int T) {
    pdl_enter();
    constexpr int H = H3, NT = 128 / UPW * 32;
    constexpr int LPU = 32 / UPW;                                 // lanes that end up with d h_prev of one unit (4 or 2)
    __shared__ __align__(16) float dgh_s[2 * 3 * H];
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const int b = blockIdx.x;
    const int lane = tid & 31;
    const int ub = (tid >> 5) * UPW;
    const float* whh = dir ? whh1 : whh0;
    float* gbih = dir ? gbih1 : gbih0;
    float* gbhh = dir ? gbhh1 : gbhh0;

    // w[i][c] = (W_hh[12 lane + 2 c][ub + i], W_hh[12 lane + 2 c + 1][ub + i])
    float2 w[UPW][6];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        const float* r0 = whh + (size_t)(12 * lane + 2 * c) * H + ub;
#pragma unroll
        for (int i4 = 0; i4 < UPW / 4; i4++) {
            const float4 a = *reinterpret_cast<const float4*>(r0 + 4 * i4);
            const float4 bq = *reinterpret_cast<const float4*>(r0 + H + 4 * i4);
            w[4 * i4][c] = make_float2(a.x, bq.x);
            w[4 * i4 + 1][c] = make_float2(a.y, bq.y);
            w[4 * i4 + 2][c] = make_float2(a.z, bq.z);
            w[4 * i4 + 3][c] = make_float2(a.w, bq.w);
        }
    }
    for (int i = tid; i < 2 * 3 * H; i += NT) dgh_s[i] = 0.f;
    float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_hn = 0.f;

    // after the butterfly: UPW = 8 -> unit = lane bits 4,3,2 ; UPW = 16 -> lane bits 4,3,2,1
    const int ul = UPW == 8 ? (lane >> 2) : (lane >> 1);
    const int sub = UPW == 8 ? (lane & 3) : (lane & 1);
    const int um = ub + ul;
    // processing order: time index t = dir ? T-1-step : step for step = T-1 .. 0
    const int t0 = dir ? 0 : T - 1;
    const ptrdiff_t ts = dir ? 1 : -1;
    const size_t bt0 = (size_t)b * T + t0;
    const float* gop = gout + bt0 * 2 * H + dir * H + um;
    const float* gsp = (dir ? gates1 : gates0) + bt0 * 4 * H + um;
    const float* hpp = (dir ? hprev1 : hprev0) + bt0 * H + um;
    float* const dgb = (dir ? dgi1 : dgi0) + bt0 * 3 * H + um;
    float* const dhb = (dir ? dghn1 : dghn0) + bt0 * H + um;
    // LPU = 4: sub 0 -> dgi r, z;  sub 1 -> dgi n, dghn;  sub 2 -> the three shared-memory values;  sub 3 -> nothing
    // LPU = 2: sub 0 -> dgi r, z + shared r, z;  sub 1 -> dgi n, dghn + shared hn
    float* pa = sub == 0 ? dgb : dgb + 2 * H;
    float* pb = sub == 0 ? dgb + H : dhb;
    const ptrdiff_t sa = ts * 3 * H, sbs = ts * (sub == 0 ? 3 * H : H);
    const bool do_g = sub < 2;
    float* dwr = dgh_s + um;                                      // + H: z, + 2 H: hn
    const float* drd = dgh_s + 12 * lane;
    float dh = 0.f;
    float p_go = gop[0], p_r = gsp[0], p_z = gsp[H], p_n = gsp[2 * H], p_ghn = gsp[3 * H], p_hp = hpp[0];
    __syncthreads();

    for (int step = T - 1; step >= 0; step--) {
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
        }
        if (LPU == 4) {
            if (sub == 2) {
                dwr[0] = dr_pre;
                dwr[H] = dz_pre;
                dwr[2 * H] = dhn;
            }
        } else {
            if (sub == 0) {
                dwr[0] = dr_pre;
                dwr[H] = dz_pre;
            } else {
                dwr[2 * H] = dhn;
            }
        }
        pa += sa;
        pb += sbs;
        sb_r += dr_pre; sb_z += dz_pre; sb_n += dn_pre; sb_hn += dhn;
        p_go = n_go; p_r = n_r; p_z = n_z; p_n = n_n; p_ghn = n_ghn; p_hp = n_hp;
        __syncthreads();
        float2 acc[UPW];
#pragma unroll
        for (int i = 0; i < UPW; i++) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float4 d4 = *reinterpret_cast<const float4*>(drd + 4 * c);
            const float2 dl = lo2(d4), dhh = hi2(d4);
#pragma unroll
            for (int i = 0; i < UPW; i++) {
                acc[i] = __ffma2_rn(w[i][2 * c], dl, acc[i]);
                acc[i] = __ffma2_rn(w[i][2 * c + 1], dhh, acc[i]);
            }
        }
        float v[UPW];
#pragma unroll
        for (int i = 0; i < UPW; i++) v[i] = acc[i].x + acc[i].y;
        float s;
        if (UPW == 16) {
            fold<16>(v, (lane & 16) != 0, 16);
            fold<8>(v, (lane & 8) != 0, 8);
            fold<4>(v, (lane & 4) != 0, 4);
            fold<2>(v, (lane & 2) != 0, 2);
            s = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
        } else {
            fold<8>(v, (lane & 16) != 0, 16);
            fold<4>(v, (lane & 8) != 0, 8);
            fold<2>(v, (lane & 4) != 0, 4);
            s = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
        }
        dh = dh_direct + s;
        // the next step's gate writes go to the other buffer: no second barrier needed
        const ptrdiff_t flip = ((T - 1 - step) & 1) ? -3 * H : 3 * H;
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

// ================================================================================================================
// H = 192 (2024 recipe): W_hh (576 x 192 fp32 = 432 KB) does not fit one SM's register file, so the hidden units are split
// over a cluster of 3 CTAs (64 units = 192 gate rows each, 144 weights per thread in registers at 256 threads).  Same
// layouts as above inside a CTA (forward: octet owns 2 units, lane = 24-wide k-slice; backward: lane = 18-row j-slice,
// warp owns 8 columns); the new h / dgh values are written into the shared memory of all three CTAs (DSMEM) and ONE
// cluster barrier per step (arrive.release ... wait.acquire, with the step's global stores in between) publishes them.
constexpr int HC = 192, CSC = 3, HUC = HC / CSC;       // 64 units per CTA
constexpr int HPADC = HC + 8;                          // padded h buffer: k -> k + 4 (k / 96)

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t map_to_rank(const void* smem_ptr, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_ptr), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
// remote store that signals the destination CTA's mbarrier when it lands (no cluster-wide fence, no L1 flush)
__device__ __forceinline__ void st_async_f32(uint32_t addr, float v, uint32_t mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n" ::"r"(addr),
                 "r"(__float_as_uint(v)), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

__global__ void __launch_bounds__(256, 1)
gru_fwd_c3_kernel(const float* __restrict__ gi0, const float* __restrict__ gi1, const float* __restrict__ whh0,
                  const float* __restrict__ whh1, const float* __restrict__ bhh0, const float* __restrict__ bhh1,
                  float* __restrict__ out, float* __restrict__ gates0, float* __restrict__ gates1,
                  float* __restrict__ hprev0, float* __restrict__ hprev1, int T, int save) {
    constexpr int H = HC, R = 6;
    __shared__ __align__(16) float h_s[2 * HPADC];
    __shared__ __align__(8) uint64_t mbar[2];     // mbar[i]: "buffer i holds the h of all three CTAs" (H * 4 bytes per phase)
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const int b = blockIdx.x / CSC;
    const uint32_t crank = cluster_rank();
    const int l8 = tid & 7;
    const int ub = (int)crank * HUC + (tid >> 3) * 2;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        mbar_expect_tx(&mbar[1], H * 4);              // step 0 publishes into buffer 1 (waited for at step 1)
        mbar_expect_tx(&mbar[0], H * 4);              // step 1 publishes into buffer 0 (waited for at step 2)
    }
    const float* whh = dir ? whh1 : whh0;
    const float* bhh = dir ? bhh1 : bhh0;

    // w[i * 3 + g][c] = W_hh[g H + ub + i][24 l8 + 2 c .. + 2]
    float2 w[R][12];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int g = 0; g < 3; g++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const float4 a = *reinterpret_cast<const float4*>(whh + (size_t)(g * H + ub + i) * H + 24 * l8 + 4 * c);
                w[i * 3 + g][2 * c] = lo2(a);
                w[i * 3 + g][2 * c + 1] = hi2(a);
            }
    for (int i = tid; i < 2 * HPADC; i += 256) h_s[i] = 0.f;
    const int ul = l8 >> 2, sub = l8 & 3;
    const int u = ub + ul;
    const float bhr = bhh[u], bhz = bhh[H + u], bhn = bhh[2 * H + u];

    const int t0 = dir ? T - 1 : 0;
    const ptrdiff_t ts = dir ? -1 : 1;
    const size_t bt0 = (size_t)b * T + t0;
    const float* gp = (dir ? gi1 : gi0) + bt0 * 3 * H + u;
    const ptrdiff_t gstep = ts * 3 * H;
    float* const obase = out + bt0 * 2 * H + dir * H + u;
    float* const gbase = (dir ? gates1 : gates0) + bt0 * 4 * H + u;
    float* const hbase = (dir ? hprev1 : hprev0) + bt0 * H + u;
    // sub 0 -> {out, hprev} + publishes h_new, sub 1 -> {r, z}, sub 2 -> {n, hn}, sub 3 -> nothing
    float* pa = sub == 0 ? obase : (sub == 1 ? gbase : gbase + 2 * H);
    float* pb = sub == 0 ? hbase : (sub == 1 ? gbase + H : gbase + 3 * H);
    const ptrdiff_t sa = ts * (sub == 0 ? 2 * H : 4 * H), sb = ts * (sub == 0 ? H : 4 * H);
    const bool do_a = sub == 0 || (save != 0 && sub < 3);
    const bool do_b = save != 0 && sub < 3;
    const float* hrd = h_s + 24 * l8 + 4 * (l8 >> 2);             // this lane's 24 h values (padded layout), buffer 0
    // where sub 0 publishes h_new: the same slot of buffer 1 in each of the three CTAs
    float* hloc = h_s + HPADC + u + 4 * (u / 96);
    uint32_t hw0 = map_to_rank(hloc, 0), hw1 = map_to_rank(hloc, 1), hw2 = map_to_rank(hloc, 2);
    uint32_t mb0 = map_to_rank(&mbar[1], 0), mb1 = map_to_rank(&mbar[1], 1), mb2 = map_to_rank(&mbar[1], 2);
    float hval = 0.f;
    float gir = gp[0] + bhr, giz = gp[H] + bhz, gin = gp[2 * H];
    // all three CTAs have zeroed their buffers and initialised their mbarriers before anyone writes remotely (the only
    // cluster-wide barrier of the kernel; per step the hand-over is st.async -> mbarrier, see below)
    cluster_arrive();
    cluster_wait();

    for (int step = 0; step < T; step++) {
        float nir = 0.f, niz = 0.f, nin = 0.f;
        if (step + 1 < T) {
            gp += gstep;
            nir = gp[0]; niz = gp[H]; nin = gp[2 * H];
        }
        if (step > 0) {
            // buffer (step & 1) is complete when the 3 x 64 st.async of step - 1 have landed.  No "buffer free" signal is
            // needed the other way: a CTA can only publish step s + 1 after it has received every CTA's step-s values,
            // which each CTA sends after its own step-s reads of the buffer being overwritten.
            mbar_wait(&mbar[step & 1], (uint32_t)((step - 1) >> 1) & 1u);
            if (tid == 0 && step + 2 < T) mbar_expect_tx(&mbar[step & 1], H * 4);     // re-arm: waited for again at step + 2
        }
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const float4 h4 = *reinterpret_cast<const float4*>(hrd + 4 * c);
            const float2 hl = lo2(h4), hh = hi2(h4);
#pragma unroll
            for (int r = 0; r < R; r++) {
                acc[r] = __ffma2_rn(w[r][2 * c], hl, acc[r]);
                acc[r] = __ffma2_rn(w[r][2 * c + 1], hh, acc[r]);
            }
        }
        float v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = acc[r].x + acc[r].y;
        fold<6>(v, (l8 & 4) != 0, 4);
        float sr = v[0], sz = v[1], sn = v[2];
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
        const float hnew = fmaf(zg, hval - n, n);
        if (sub == 0 && step + 1 < T) {                          // nobody reads the h of the last step from shared memory
            st_async_f32(hw0, hnew, mb0);
            st_async_f32(hw1, hnew, mb1);
            st_async_f32(hw2, hnew, mb2);
        }
        if (do_a) *pa = sub == 0 ? hnew : (sub == 1 ? r : n);
        if (do_b) *pb = sub == 0 ? hval : (sub == 1 ? zg : ghn);
        pa += sa;
        pb += sb;
        hval = hnew;
        gir = nir + bhr; giz = niz + bhz; gin = nin;
        const int flip = (step & 1) ? -HPADC : HPADC;             // readers move to the buffer just written
        const int mflip = (step & 1) ? 8 : -8;                    // mbar[1] -> mbar[0] -> mbar[1] ... (byte addresses)
        hrd += flip;
        hw0 -= 4 * flip; hw1 -= 4 * flip; hw2 -= 4 * flip;        // byte addresses
        mb0 += mflip; mb1 += mflip; mb2 += mflip;
    }
    // every st.async aimed at this CTA has been waited for (the last step publishes nothing): safe to exit
}

__global__ void __launch_bounds__(256, 1)
gru_bwd_c3_kernel(const float* __restrict__ gout, const float* __restrict__ whh0, const float* __restrict__ whh1,
                  const float* __restrict__ gates0, const float* __restrict__ gates1, const float* __restrict__ hprev0,
                  const float* __restrict__ hprev1, float* __restrict__ dgi0, float* __restrict__ dgi1,
                  float* __restrict__ dghn0, float* __restrict__ dghn1, float* __restrict__ gbih0,
                  float* __restrict__ gbih1, float* __restrict__ gbhh0, float* __restrict__ gbhh1, int T) {
    constexpr int H = HC, UPW = 8;
    __shared__ __align__(16) float dgh_s[2 * 3 * H];
    __shared__ __align__(8) uint64_t mbar[2];     // mbar[i]: "buffer i holds d(r, z, hn) of all three CTAs" (3 H * 4 bytes)
    const int tid = threadIdx.x;
    const int dir = blockIdx.y;
    const int b = blockIdx.x / CSC;
    const uint32_t crank = cluster_rank();
    const int lane = tid & 31;
    const int ub = (int)crank * HUC + (tid >> 5) * UPW;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        mbar_expect_tx(&mbar[0], 3 * H * 4);
        mbar_expect_tx(&mbar[1], 3 * H * 4);
    }
    const float* whh = dir ? whh1 : whh0;
    float* gbih = dir ? gbih1 : gbih0;
    float* gbhh = dir ? gbhh1 : gbhh0;

    // w[i][c] = (W_hh[18 lane + 2 c][ub + i], W_hh[18 lane + 2 c + 1][ub + i])
    float2 w[UPW][9];
#pragma unroll
    for (int c = 0; c < 9; c++) {
        const float* r0 = whh + (size_t)(18 * lane + 2 * c) * H + ub;
#pragma unroll
        for (int i4 = 0; i4 < UPW / 4; i4++) {
            const float4 a = *reinterpret_cast<const float4*>(r0 + 4 * i4);
            const float4 bq = *reinterpret_cast<const float4*>(r0 + H + 4 * i4);
            w[4 * i4][c] = make_float2(a.x, bq.x);
            w[4 * i4 + 1][c] = make_float2(a.y, bq.y);
            w[4 * i4 + 2][c] = make_float2(a.z, bq.z);
            w[4 * i4 + 3][c] = make_float2(a.w, bq.w);
        }
    }
    for (int i = tid; i < 2 * 3 * H; i += 256) dgh_s[i] = 0.f;
    float sb_r = 0.f, sb_z = 0.f, sb_n = 0.f, sb_hn = 0.f;
    const int ul = lane >> 2, sub = lane & 3;
    const int um = ub + ul;
    const int t0 = dir ? 0 : T - 1;
    const ptrdiff_t ts = dir ? 1 : -1;
    const size_t bt0 = (size_t)b * T + t0;
    const float* gop = gout + bt0 * 2 * H + dir * H + um;
    const float* gsp = (dir ? gates1 : gates0) + bt0 * 4 * H + um;
    const float* hpp = (dir ? hprev1 : hprev0) + bt0 * H + um;
    float* const dgb = (dir ? dgi1 : dgi0) + bt0 * 3 * H + um;
    float* const dhb = (dir ? dghn1 : dghn0) + bt0 * H + um;
    // sub 0 -> dgi r, z;  sub 1 -> dgi n, dghn;  sub 2 -> publishes (d r_pre, d z_pre, d hn) to the three CTAs
    float* pa = sub == 0 ? dgb : dgb + 2 * H;
    float* pb = sub == 0 ? dgb + H : dhb;
    const ptrdiff_t sa = ts * 3 * H, sbs = ts * (sub == 0 ? 3 * H : H);
    const bool do_g = sub < 2;
    uint32_t dw0 = map_to_rank(dgh_s + um, 0), dw1 = map_to_rank(dgh_s + um, 1), dw2 = map_to_rank(dgh_s + um, 2);
    uint32_t mb0 = map_to_rank(&mbar[0], 0), mb1 = map_to_rank(&mbar[0], 1), mb2 = map_to_rank(&mbar[0], 2);
    const float* drd = dgh_s + 18 * lane;
    float dh = 0.f;
    float p_go = gop[0], p_r = gsp[0], p_z = gsp[H], p_n = gsp[2 * H], p_ghn = gsp[3 * H], p_hp = hpp[0];
    cluster_arrive();
    cluster_wait();

    for (int step = T - 1; step >= 0; step--) {
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
        if (sub == 2) {
            st_async_f32(dw0, dr_pre, mb0); st_async_f32(dw0 + 4 * H, dz_pre, mb0); st_async_f32(dw0 + 8 * H, dhn, mb0);
            st_async_f32(dw1, dr_pre, mb1); st_async_f32(dw1 + 4 * H, dz_pre, mb1); st_async_f32(dw1 + 8 * H, dhn, mb1);
            st_async_f32(dw2, dr_pre, mb2); st_async_f32(dw2 + 4 * H, dz_pre, mb2); st_async_f32(dw2 + 8 * H, dhn, mb2);
        }
        if (do_g) {
            *pa = sub == 0 ? dr_pre : dn_pre;
            *pb = sub == 0 ? dz_pre : dhn;
        }
        pa += sa;
        pb += sbs;
        sb_r += dr_pre; sb_z += dz_pre; sb_n += dn_pre; sb_hn += dhn;
        p_go = n_go; p_r = n_r; p_z = n_z; p_n = n_n; p_ghn = n_ghn; p_hp = n_hp;
        {
            // processed-step counter k = T - 1 - step: buffer k & 1, phase parity (k >> 1) & 1; re-armed for step k + 2.
            // (No "buffer free" signal needed: a CTA publishes step k + 2 only after it has consumed every CTA's step k + 1
            // values, which each CTA sends after its own step-k reads of this buffer.)
            const int k = T - 1 - step;
            mbar_wait(&mbar[k & 1], (uint32_t)(k >> 1) & 1u);
            if (tid == 0 && k + 2 < T) mbar_expect_tx(&mbar[k & 1], 3 * H * 4);
        }
        float2 acc[UPW];
#pragma unroll
        for (int i = 0; i < UPW; i++) acc[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 9; c++) {
            const float2 d2 = *reinterpret_cast<const float2*>(drd + 2 * c);
#pragma unroll
            for (int i = 0; i < UPW; i++) acc[i] = __ffma2_rn(w[i][c], d2, acc[i]);
        }
        float v[UPW];
#pragma unroll
        for (int i = 0; i < UPW; i++) v[i] = acc[i].x + acc[i].y;
        fold<8>(v, (lane & 16) != 0, 16);
        fold<4>(v, (lane & 8) != 0, 8);
        fold<2>(v, (lane & 4) != 0, 4);
        float s = v[0] + __shfl_xor_sync(0xffffffffu, v[0], 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        dh = dh_direct + s;
        const int flip = ((T - 1 - step) & 1) ? -3 * H : 3 * H;   // the next step's writes go to the other buffer
        const int mflip = ((T - 1 - step) & 1) ? -8 : 8;
        drd += flip;
        dw0 += 4 * flip; dw1 += 4 * flip; dw2 += 4 * flip;
        mb0 += mflip; mb1 += mflip; mb2 += mflip;
    }
    if (sub == 0 && gbih != nullptr) {
        atomicAdd(&gbih[um], sb_r);
        atomicAdd(&gbih[H + um], sb_z);
        atomicAdd(&gbih[2 * H + um], sb_n);
        atomicAdd(&gbhh[um], sb_r);
        atomicAdd(&gbhh[H + um], sb_z);
        atomicAdd(&gbhh[2 * H + um], sb_hn);
    }
    // no CTA may exit while a sibling can still write into its shared memory
    cluster_arrive();
    cluster_wait();
}

template <class Kern, class... Args>
int launch_cluster3(Kern kern, int B, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * CSC, 2);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CSC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SEDK_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
    count_launch();
    return SEDK_OK;
}

template <int UPO>
int run_fwd_v3(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
               float* const gates[2], float* const hprev[2], int B, int T, int save, cudaStream_t s) {
    SEDK_CUDA(pdl_launch(gru_fwd_v3_kernel<UPO>, dim3(dim3(B, 2)), dim3(128 / UPO * 8), (size_t)(0), s, gi[0], gi[1], w_hh[0], w_hh[1], b_hh[0], b_hh[1], out,
                                                               gates[0], gates[1], hprev[0], hprev[1], T, save));
    SEDK_LAUNCH_CHECK("gru_fwd_v3_kernel");
    return SEDK_OK;
}

template <int UPW>
int run_bwd_v3(const float* gout, const float* const w_hh[2], const float* const gates[2], const float* const hprev[2],
               float* const dgi[2], float* const dghn[2], float* const gb_ih[2], float* const gb_hh[2], int B, int T,
               int zeroed, cudaStream_t s) {
    for (int d = 0; d < 2 && !zeroed; d++) {
        SEDK_CUDA(cudaMemsetAsync(gb_ih[d], 0, (size_t)3 * H3 * sizeof(float), s));
        SEDK_CUDA(cudaMemsetAsync(gb_hh[d], 0, (size_t)3 * H3 * sizeof(float), s));
    }
    SEDK_CUDA(pdl_launch(gru_bwd_v3_kernel<UPW>, dim3(dim3(B, 2)), dim3(128 / UPW * 32), (size_t)(0), s, gout, w_hh[0], w_hh[1], gates[0], gates[1], hprev[0],
                                                                hprev[1], dgi[0], dgi[1], dghn[0], dghn[1], gb_ih[0],
                                                                gb_ih[1], gb_hh[0], gb_hh[1], T));
    SEDK_LAUNCH_CHECK("gru_bwd_v3_kernel");
    return SEDK_OK;
}

}  // namespace

// variant 1: 8 warps, every weight in registers ("fat" threads); variant 2: 16 warps (half the weights per thread)
int launch_gru_fwd_v3(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
                      float* const gates[2], float* const hprev[2], int B, int T, int save, int variant, cudaStream_t s) {
    if (variant == 2) return run_fwd_v3<2>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
    return run_fwd_v3<4>(gi, w_hh, b_hh, out, gates, hprev, B, T, save, s);
}

int launch_gru_bwd_v3(const float* gout, const float* const w_hh[2], const float* const gates[2],
                      const float* const hprev[2], float* const dgi[2], float* const dghn[2], float* const gb_ih[2],
                      float* const gb_hh[2], int B, int T, int zeroed, int variant, cudaStream_t s) {
    if (variant == 2) return run_bwd_v3<8>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
    return run_bwd_v3<16>(gout, w_hh, gates, hprev, dgi, dghn, gb_ih, gb_hh, B, T, zeroed, s);
}

int launch_gru_fwd_c3(const float* const gi[2], const float* const w_hh[2], const float* const b_hh[2], float* out,
                      float* const gates[2], float* const hprev[2], int B, int T, int save, cudaStream_t s) {
    return launch_cluster3(gru_fwd_c3_kernel, B, s, gi[0], gi[1], w_hh[0], w_hh[1], b_hh[0], b_hh[1], out, gates[0], gates[1],
                           hprev[0], hprev[1], T, save);
}

int launch_gru_bwd_c3(const float* gout, const float* const w_hh[2], const float* const gates[2],
                      const float* const hprev[2], float* const dgi[2], float* const dghn[2], float* const gb_ih[2],
                      float* const gb_hh[2], int B, int T, int zeroed, cudaStream_t s) {
    for (int d = 0; d < 2 && !zeroed; d++) {
        SEDK_CUDA(cudaMemsetAsync(gb_ih[d], 0, (size_t)3 * HC * sizeof(float), s));
        SEDK_CUDA(cudaMemsetAsync(gb_hh[d], 0, (size_t)3 * HC * sizeof(float), s));
    }
    return launch_cluster3(gru_bwd_c3_kernel, B, s, gout, w_hh[0], w_hh[1], gates[0], gates[1], hprev[0], hprev[1], dgi[0],
                           dgi[1], dghn[0], dghn[1], gb_ih[0], gb_ih[1], gb_hh[0], gb_hh[1], T);
}

}  // namespace sedk
