// Fused waveform -> |STFT| -> mel -> (log) kernel for sm_100a, second generation (same decomposition as logmel.cu, which
// stays in the library as the A/B baseline behind option "logmel_v2" = 0).  What changed, from the ncu profile of the first
// generation (168 registers and 94 KB of shared memory per 4-warp CTA -> 8 warps per SM, issue slots 34 % busy, 3 460
// warp-instructions per frame):
//   * 4 CTAs (16 warps) per SM: 128 registers per thread (no spills), and the Hamming window, the W_1024 step twiddles and
//     the W_2048 untangle twiddles are read through L1 (__ldg, coalesced, 24 KB shared by every warp of the SM) instead
//     of living in every CTA's shared memory; the magnitude buffer aliases the FFT exchange buffer -> 53 KB per CTA;
//   * packed fp32 arithmetic (FADD2 / FFMA2 / FMUL2) for the 160 butterflies of the two in-register 32-point FFTs and the
//     window multiply: half the issue slots for the same flops;
//   * the real-FFT untangle works on conjugate pairs: X[k] and X[1024 - k] share Xe[k] and T = W^k Xo[k]
//     (|X[k]| = |Xe + T|, |X[1024 - k]| = |Xe - T|), 16 pair-iterations per lane instead of 32 single bins, magnitudes
//     through rsqrt.approx (2 ulp; the log-mel budget is 1.2e-5 relative).
//
// Replaces torchaudio MelSpectrogram(n_fft=win=2048, hop, center=True/reflect, hamming, power=1, HTK fb) and
// AmplitudeToDB + clamp as used at recipes/dcase2023_task4_baseline/local/sed_trainer.py:79-91,253-264,282.
// The 1025 x T spectrum never touches HBM: per clip the kernel reads 4*L bytes and writes 4*n_mels*T bytes.
//
// Work decomposition
//   CTA   : FR = 8 consecutive frames of one clip (persistent over frame groups); the (FR-1)*hop + 2048 samples
//           they share are staged ONCE into shared memory with a 1-D bulk async copy (TMA engine, UBLKCP).
//   warp  : one frame at a time.  real 2048-FFT = complex 1024-FFT of z[n] = x[2n] + i x[2n+1] (windowed), done
//           as 32 x 32: every lane runs a 32-point FFT in registers, one padded smem transpose with the
//           W_1024^{n2 k1} twiddles, a second 32-point FFT in registers, then the real-FFT untangle + magnitude.
//   mel   : each lane owns 4 triangular filters (lane, lane+32, ...) and walks their non-zero runs over the
//           magnitudes in shared memory (sparse filterbank: 2 024 of 131 200 weights are non-zero).
#include "common.cuh"
#include <type_traits>

namespace sedk {
namespace {
namespace v2 {

constexpr int kNfft = 2048;
constexpr int kHalf = 1024;
constexpr int FR = 8;   // frames per CTA group
constexpr int NW = 4;   // warps per CTA
constexpr int SROW = 33;

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__host__ __device__ constexpr float cos32(int m) {
    // cos(2 pi m / 32), m in [0, 16]
    switch (m) {
        case 0: return 1.0f;
        case 1: return 0.98078528040323044913f;
        case 2: return 0.92387953251128675613f;
        case 3: return 0.83146961230254523708f;
        case 4: return 0.70710678118654752440f;
        case 5: return 0.55557023301960222474f;
        case 6: return 0.38268343236508977173f;
        case 7: return 0.19509032201612826785f;
        case 8: return 0.0f;
        case 9: return -0.19509032201612826785f;
        case 10: return -0.38268343236508977173f;
        case 11: return -0.55557023301960222474f;
        case 12: return -0.70710678118654752440f;
        case 13: return -0.83146961230254523708f;
        case 14: return -0.92387953251128675613f;
        case 15: return -0.98078528040323044913f;
        default: return -1.0f;
    }
}
__host__ __device__ constexpr float sin32(int m) { return m <= 8 ? cos32(8 - m) : cos32(m - 8); }

// v * W_32^M,  W_32 = exp(-2 pi i / 32)
template <int M>
__device__ __forceinline__ float2 mul_w32(float2 v) {
    if constexpr (M == 0) {
        return v;
    } else if constexpr (M == 8) {
        return make_float2(v.y, -v.x);
    } else if constexpr (M == 4) {
        constexpr float r = cos32(4);
        return make_float2((v.x + v.y) * r, (v.y - v.x) * r);
    } else if constexpr (M == 12) {
        constexpr float r = cos32(4);
        return make_float2((v.y - v.x) * r, -(v.x + v.y) * r);
    } else {
        constexpr float c = cos32(M), s = sin32(M);
        return make_float2(fmaf(v.y, s, v.x * c), fmaf(v.y, c, -(v.x * s)));
    }
}

template <int S>
__device__ __forceinline__ void dif_stage(float2 (&v)[32]) {
    static_for<0, 32 / (2 * S)>([&](auto blk) {
        static_for<0, S>([&](auto jj) {
            constexpr int base = decltype(blk)::value * 2 * S;
            constexpr int j = decltype(jj)::value;
            constexpr int m = j * (16 / S);
            const float2 a = v[base + j], b = v[base + j + S];
            v[base + j] = __fadd2_rn(a, b);
            v[base + j + S] = mul_w32<m>(__ffma2_rn(b, make_float2(-1.0f, -1.0f), a));
        });
    });
}

// in-register 32-point forward DFT; on return v[i] = X[bitrev5(i)]
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
    dif_stage<16>(v);
    dif_stage<8>(v);
    dif_stage<4>(v);
    dif_stage<2>(v);
    dif_stage<1>(v);
}
__host__ __device__ constexpr int bitrev5(int i) {
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

__device__ __forceinline__ int reflect_idx(int i, int L) {
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    return i;
}

struct SmemLayout {
    int chunk_floats;
    size_t off_chunk, off_tile, off_warp, off_bar, total;
};
__host__ __device__ inline SmemLayout make_layout(int hop, int n_mels) {
    SmemLayout s;
    s.chunk_floats = ((FR - 1) * hop + kNfft + 8 + 3) & ~3;
    size_t o = 0;
    s.off_bar = o;     o += 16;
    s.off_chunk = o;   o += (size_t)s.chunk_floats * 4;
    s.off_tile = o;    o += (size_t)FR * n_mels * 4;
    o = (o + 15) & ~(size_t)15;
    s.off_warp = o;    o += (size_t)NW * (32 * SROW * 8);
    s.total = o;
    return s;
}

// PCM sample types: fp32 waveforms, or the int16 the datasets are stored in (f2, input pipeline): torchaudio.load hands the
// reference x / 32768 as fp32 (an exact power-of-two scale), so converting in the load path is bit-identical and halves
// both the H2D copy and the HBM read (320 KB instead of 640 KB per clip).
__device__ __forceinline__ float pcm(float v) { return v; }
__device__ __forceinline__ float pcm(int16_t v) { return (float)v * (1.0f / 32768.0f); }
__device__ __forceinline__ float2 pcm_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 pcm_pair(const int16_t* p) {
    const short2 q = *reinterpret_cast<const short2*>(p);
    return make_float2((float)q.x * (1.0f / 32768.0f), (float)q.y * (1.0f / 32768.0f));
}

template <typename TS>
__global__ void __launch_bounds__(NW * 32, 4)
logmel2_kernel(const TS* __restrict__ wave, int B, int L, int T, sedk_mel_tables tab, float* __restrict__ out,
              int64_t out_sb, int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi,
              uint32_t* __restrict__ minmax, int n_groups_per_clip) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int hop = tab.hop, n_mels = tab.n_mels;
    const SmemLayout lay = make_layout(hop, n_mels);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.off_bar);
    const float2* g_window2 = reinterpret_cast<const float2*>(tab.window);      // read through L1 (coalesced)
    const float2* g_tw2048 = reinterpret_cast<const float2*>(tab.tw2048);
    const float2* g_tw32 = reinterpret_cast<const float2*>(tab.tw32x32);
    TS* s_chunk = reinterpret_cast<TS*>(smem + lay.off_chunk);
    constexpr int AL = 16 / (int)sizeof(TS);          // samples per 16-byte unit of the bulk copy
    float* s_tile = reinterpret_cast<float*>(smem + lay.off_tile);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem + lay.off_warp + (size_t)warp * (32 * SROW * 8);
    float2* s_x = reinterpret_cast<float2*>(wbase);                   // 32 x 33 exchange, later Z[1024]
    float* s_mag = reinterpret_cast<float*>(wbase);                   // |X[k]|, k = 0..1024: aliases s_x once Z is in registers

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int n_groups = B * n_groups_per_clip;
    uint32_t phase = 0;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int b = grp / n_groups_per_clip;
        const int f0 = (grp - b * n_groups_per_clip) * FR;
        const int nfr = min(FR, T - f0);
        const TS* clip = wave + (size_t)b * L;
        // samples [c0, c1) of this clip cover every (reflected) index the group's frames touch
        int c0 = f0 * hop - kHalf - AL;
        c0 = c0 < 0 ? 0 : (c0 & ~(AL - 1));
        int c1 = min(L, (f0 + nfr - 1) * hop + kHalf);
        // capacity of the staging buffer in samples (int16 samples take half the bytes); cannot be exceeded for hop % 4 == 0,
        // the clamp guards odd hops
        const int cap = lay.chunk_floats * (4 / (int)sizeof(TS));
        if (c1 - c0 > cap) c1 = c0 + cap;
        const int n_chunk = c1 - c0;
        const TS* src = clip + c0;
        const bool bulk_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const int n_bulk = bulk_ok ? (n_chunk & ~(AL - 1)) : 0;
        if (n_bulk > 0 && threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, (uint32_t)n_bulk * (uint32_t)sizeof(TS));
            bulk_g2s(s_chunk, src, (uint32_t)n_bulk * (uint32_t)sizeof(TS), bar);
        }
        for (int i = n_bulk + threadIdx.x; i < n_chunk; i += blockDim.x) s_chunk[i] = src[i];
        if (n_bulk > 0) {
            mbar_wait(bar, phase);
            phase ^= 1;
        }
        __syncthreads();

        for (int fl = warp; fl < nfr; fl += NW) {
            const int t = f0 + fl;
            const int s0 = t * hop - kHalf;   // first sample of the frame (may be negative)
            float2 v[32];
            const bool interior = (s0 >= 0) && (s0 + kNfft <= L) && (((s0 - c0) & 1) == 0);
            if (interior) {
                const TS* xs = s_chunk + (s0 - c0);
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++)
                    v[n1] = __fmul2_rn(pcm_pair(xs + 2 * (32 * n1 + lane)), __ldg(g_window2 + 32 * n1 + lane));
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    int j = 2 * (32 * n1 + lane);
                    int i0 = reflect_idx(s0 + j, L) - c0, i1 = reflect_idx(s0 + j + 1, L) - c0;
                    i0 = min(max(i0, 0), n_chunk - 1);
                    i1 = min(max(i1, 0), n_chunk - 1);
                    const float2 w = __ldg(g_window2 + (j >> 1));
                    v[n1] = make_float2(pcm(s_chunk[i0]) * w.x, pcm(s_chunk[i1]) * w.y);
                }
            }
            // ---- step 1: lane = n2, FFT over n1; twiddle W_1024^{n2 k1}; transpose through smem
            fft32(v);
            static_for<0, 32>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                constexpr int k1 = bitrev5(i);
                const float2 w = __ldg(g_tw32 + k1 * 32 + lane);
                float2 a = v[i];
                s_x[k1 * SROW + lane] = make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
            });
            __syncwarp();
            // ---- step 2: lane = k1, FFT over n2 -> Z[k1 + 32 k2]
#pragma unroll
            for (int n2 = 0; n2 < 32; n2++) v[n2] = s_x[lane * SROW + n2];
            __syncwarp();
            fft32(v);
            static_for<0, 32>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                constexpr int k2 = bitrev5(i);
                s_x[lane + 32 * k2] = v[i];
            });
            __syncwarp();
            // ---- untangle on conjugate pairs.  With A = Z[k], B = Z[1024 - k] (k = 1..511):
            //   2 Xe[k] = A + conj(B),  2 Xo[k] = -i (A - conj(B)),  T = W_2048^k Xo[k]
            //   X[k] = Xe + T,  X[1024 - k] = conj(Xe - T)   ->   |X[k]| = |Xe + T|, |X[1024 - k]| = |Xe - T|
            // Every Z value a lane needs is read into registers first, so the magnitudes can overwrite the Z buffer.
            float2 za[16], zb[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int k = lane + 32 * j;
                za[j] = s_x[k];
                zb[j] = s_x[(kHalf - k) & (kHalf - 1)];
            }
            const float2 z512 = s_x[512];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int k = lane + 32 * j;
                const float2 a = za[j], bb = zb[j];
                const float2 w = __ldg(g_tw2048 + k);
                const float er = a.x + bb.x, ei = a.y - bb.y;          // 2 Xe
                const float orr = a.y + bb.y, oi = bb.x - a.x;         // 2 Xo
                const float tr = w.x * orr - w.y * oi, ti = w.x * oi + w.y * orr;      // 2 T
                const float pr = er + tr, pi = ei + ti, mr = er - tr, mi = ei - ti;
                const float sp = pr * pr + pi * pi, sm = mr * mr + mi * mi;
                const float mp = sp > 0.f ? 0.5f * sp * rsqrtf(sp) : 0.f;
                const float mm = sm > 0.f ? 0.5f * sm * rsqrtf(sm) : 0.f;
                if (k == 0) {
                    // Z[0]: X[0] = Re + Im, X[1024] = Re - Im (both purely real)
                    s_mag[0] = fabsf(a.x + a.y);
                    s_mag[kHalf] = fabsf(a.x - a.y);
                } else {
                    s_mag[k] = mp;
                    s_mag[kHalf - k] = mm;
                }
            }
            if (lane == 0) s_mag[512] = sqrtf(z512.x * z512.x + z512.y * z512.y);      // self-paired bin: X[512] = conj(Z[512])
            __syncwarp();
            // ---- mel triangles
            for (int m = lane; m < n_mels; m += 32) {
                const int st = __ldg(tab.fb_start + m), ln = __ldg(tab.fb_len + m), of = __ldg(tab.fb_off + m);
                float acc = 0.f;
                for (int j = 0; j < ln; j++) acc = fmaf(__ldg(tab.fb_w + of + j), s_mag[st + j], acc);
                float val = acc;
                if (log_mode) {
                    val = 20.0f * log10f(fmaxf(acc, amin));
                    val = fminf(fmaxf(val, db_lo), db_hi);
                }
                s_tile[fl * n_mels + m] = val;
            }
            __syncwarp();
        }
        __syncthreads();
        // ---- write the [nfr, n_mels] tile and fold it into the clip's min / max
        float vmin = INFINITY, vmax = -INFINITY;
        float* ob = out + (size_t)b * out_sb;
        const int n_el = nfr * n_mels;
        if (out_st == 1) {
            for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
                int m = i / nfr, fl = i - m * nfr;
                float val = s_tile[fl * n_mels + m];
                ob[(int64_t)m * out_sm + (f0 + fl)] = val;
                vmin = fminf(vmin, val);
                vmax = fmaxf(vmax, val);
            }
        } else {
            for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
                int fl = i / n_mels, m = i - fl * n_mels;
                float val = s_tile[i];
                ob[(int64_t)m * out_sm + (int64_t)(f0 + fl) * out_st] = val;
                vmin = fminf(vmin, val);
                vmax = fmaxf(vmax, val);
            }
        }
        if (minmax != nullptr) {
            vmin = warp_min(vmin);
            vmax = warp_max(vmax);
            if (lane == 0) {
                atomicMin(minmax + 2 * b, f2ord(vmin));
                atomicMax(minmax + 2 * b + 1, f2ord(vmax));
            }
        }
        __syncthreads();
    }
}

}  // namespace v2
}  // namespace

template <typename TS>
static int launch_logmel_v2_t(const TS* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb,
                              int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi,
                              uint32_t* minmax, cudaStream_t stream) {
    using namespace v2;
    const int T = 1 + L / tab->hop;
    const int groups = cdiv(T, FR);
    SmemLayout lay = make_layout(tab->hop, tab->n_mels);
    SEDK_REQUIRE(lay.total <= 227 * 1024, "sedk_logmel_fwd: hop %d needs %zu B of shared memory", tab->hop, lay.total);
    auto kern = logmel2_kernel<TS>;
    int rc = opt_in_smem(kern, lay.total);
    if (rc != SEDK_OK) return rc;
    static int per_sm = 0;
    if (per_sm == 0) {
        int o = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, NW * 32, lay.total);
        per_sm = o < 1 ? 1 : o;
    }
    long long total = (long long)B * groups;
    int grid = (int)(total < (long long)num_sms() * per_sm ? total : (long long)num_sms() * per_sm);
    kern<<<grid, NW * 32, lay.total, stream>>>(wave, B, L, T, *tab, out, out_sb, out_sm, out_st, log_mode, amin, db_lo, db_hi,
                                              minmax, groups);
    SEDK_LAUNCH_CHECK("logmel2_kernel");
    return SEDK_OK;
}

int launch_logmel_v2(const float* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb, int64_t out_sm,
                     int64_t out_st, int log_mode, float amin, float db_lo, float db_hi, uint32_t* minmax, cudaStream_t stream) {
    return launch_logmel_v2_t<float>(wave, B, L, tab, out, out_sb, out_sm, out_st, log_mode, amin, db_lo, db_hi, minmax, stream);
}

int launch_logmel_v2_i16(const int16_t* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb,
                         int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi, uint32_t* minmax,
                         cudaStream_t stream) {
    return launch_logmel_v2_t<int16_t>(wave, B, L, tab, out, out_sb, out_sm, out_st, log_mode, amin, db_lo, db_hi, minmax,
                                       stream);
}

}  // namespace sedk
