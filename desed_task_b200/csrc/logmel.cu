// Fused waveform -> |STFT| -> mel -> (log) kernel for sm_100a.
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
            float2 a = v[base + j], b = v[base + j + S];
            v[base + j] = make_float2(a.x + b.x, a.y + b.y);
            v[base + j + S] = mul_w32<m>(make_float2(a.x - b.x, a.y - b.y));
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
    size_t off_window, off_tw2048, off_tw32, off_chunk, off_tile, off_warp, off_bar, total;
};
__host__ __device__ inline SmemLayout make_layout(int hop, int n_mels) {
    SmemLayout s;
    s.chunk_floats = ((FR - 1) * hop + kNfft + 8 + 3) & ~3;
    size_t o = 0;
    s.off_bar = o;     o += 16;
    s.off_window = o;  o += kNfft * 4;
    s.off_tw2048 = o;  o += kHalf * 8;
    s.off_tw32 = o;    o += kHalf * 8;
    s.off_chunk = o;   o += (size_t)s.chunk_floats * 4;
    s.off_tile = o;    o += (size_t)FR * n_mels * 4;
    o = (o + 15) & ~(size_t)15;
    s.off_warp = o;    o += (size_t)NW * (32 * SROW * 8 + 1028 * 4);
    s.total = o;
    return s;
}

__global__ void __launch_bounds__(NW * 32, 2)
logmel_kernel(const float* __restrict__ wave, int B, int L, int T, sedk_mel_tables tab, float* __restrict__ out,
              int64_t out_sb, int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi,
              uint32_t* __restrict__ minmax, int n_groups_per_clip) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int hop = tab.hop, n_mels = tab.n_mels;
    const SmemLayout lay = make_layout(hop, n_mels);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.off_bar);
    float* s_window = reinterpret_cast<float*>(smem + lay.off_window);
    float2* s_tw2048 = reinterpret_cast<float2*>(smem + lay.off_tw2048);
    float2* s_tw32 = reinterpret_cast<float2*>(smem + lay.off_tw32);
    float* s_chunk = reinterpret_cast<float*>(smem + lay.off_chunk);
    float* s_tile = reinterpret_cast<float*>(smem + lay.off_tile);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem + lay.off_warp + (size_t)warp * (32 * SROW * 8 + 1028 * 4);
    float2* s_x = reinterpret_cast<float2*>(wbase);                   // 32 x 33 exchange, later Z[1024]
    float* s_mag = reinterpret_cast<float*>(wbase + 32 * SROW * 8);   // |X[k]|, k = 0..1024

    for (int i = threadIdx.x; i < kNfft; i += blockDim.x) s_window[i] = tab.window[i];
    for (int i = threadIdx.x; i < kHalf; i += blockDim.x) {
        s_tw2048[i] = reinterpret_cast<const float2*>(tab.tw2048)[i];
        s_tw32[i] = reinterpret_cast<const float2*>(tab.tw32x32)[i];
    }
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
        const float* clip = wave + (size_t)b * L;
        // samples [c0, c1) of this clip cover every (reflected) index the group's frames touch
        int c0 = f0 * hop - kHalf - 4;
        c0 = c0 < 0 ? 0 : (c0 & ~3);
        int c1 = min(L, (f0 + nfr - 1) * hop + kHalf);
        if (c1 - c0 > lay.chunk_floats) c1 = c0 + lay.chunk_floats;   // cannot happen for hop % 4 == 0; guards odd hops
        const int n_chunk = c1 - c0;
        const float* src = clip + c0;
        const bool bulk_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const int n_bulk = bulk_ok ? (n_chunk & ~3) : 0;
        if (n_bulk > 0 && threadIdx.x == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, (uint32_t)n_bulk * 4u);
            bulk_g2s(s_chunk, src, (uint32_t)n_bulk * 4u, bar);
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
                const float2* xs = reinterpret_cast<const float2*>(s_chunk + (s0 - c0));
                const float2* ws = reinterpret_cast<const float2*>(s_window);
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    float2 x = xs[32 * n1 + lane], w = ws[32 * n1 + lane];
                    v[n1] = make_float2(x.x * w.x, x.y * w.y);
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < 32; n1++) {
                    int j = 2 * (32 * n1 + lane);
                    int i0 = reflect_idx(s0 + j, L) - c0, i1 = reflect_idx(s0 + j + 1, L) - c0;
                    i0 = min(max(i0, 0), n_chunk - 1);
                    i1 = min(max(i1, 0), n_chunk - 1);
                    v[n1] = make_float2(s_chunk[i0] * s_window[j], s_chunk[i1] * s_window[j + 1]);
                }
            }
            // ---- step 1: lane = n2, FFT over n1; twiddle W_1024^{n2 k1}; transpose through smem
            fft32(v);
            static_for<0, 32>([&](auto ii) {
                constexpr int i = decltype(ii)::value;
                constexpr int k1 = bitrev5(i);
                float2 w = s_tw32[k1 * 32 + lane];
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
            // ---- untangle: X[k] = Xe[k] + W_2048^k Xo[k], k = 0..1024; store |X[k]|
#pragma unroll 4
            for (int j = 0; j < 32; j++) {
                const int k = lane + 32 * j;
                float2 a = s_x[k], bb = s_x[(kHalf - k) & (kHalf - 1)];
                float2 w = s_tw2048[k];
                float xer = 0.5f * (a.x + bb.x), xei = 0.5f * (a.y - bb.y);
                float xor_ = 0.5f * (a.y + bb.y), xoi = -0.5f * (a.x - bb.x);
                float re = xer + (w.x * xor_ - w.y * xoi);
                float im = xei + (w.x * xoi + w.y * xor_);
                s_mag[k] = sqrtf(re * re + im * im);
            }
            if (lane == 0) {
                float2 a = s_x[0];
                // k = 1024: Xe[0] - Xo[0] = Re Z[0] - Im Z[0] (purely real)
                s_mag[kHalf] = fabsf(a.x - a.y);
            }
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

}  // namespace
}  // namespace sedk

namespace sedk {
int launch_logmel_v2(const float* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb, int64_t out_sm,
                     int64_t out_st, int log_mode, float amin, float db_lo, float db_hi, uint32_t* minmax, cudaStream_t stream);
int launch_logmel_v2_i16(const int16_t* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb,
                         int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi, uint32_t* minmax,
                         cudaStream_t stream);
}

extern "C" int sedk_logmel_fwd_i16(const int16_t* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb,
                                   int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi,
                                   uint32_t* minmax, void* stream) {
    using namespace sedk;
    SEDK_PROF("logmel", (cudaStream_t)stream);
    SEDK_REQUIRE(wave && tab && out, "sedk_logmel_fwd_i16: null pointer");
    SEDK_REQUIRE(B > 0, "sedk_logmel_fwd_i16: B must be positive (got %d)", B);
    SEDK_REQUIRE(L > kHalf, "sedk_logmel_fwd_i16: reflect padding needs L > %d samples (got %d)", kHalf, L);
    SEDK_REQUIRE(tab->hop > 0 && tab->hop <= kNfft, "sedk_logmel_fwd_i16: hop %d out of range", tab->hop);
    SEDK_REQUIRE(tab->n_mels > 0 && tab->n_mels <= 256, "sedk_logmel_fwd_i16: n_mels %d out of range", tab->n_mels);
    return launch_logmel_v2_i16(wave, B, L, tab, out, out_sb, out_sm, out_st, log_mode, amin, db_lo, db_hi, minmax,
                                (cudaStream_t)stream);
}

extern "C" int sedk_logmel_fwd(const float* wave, int B, int L, const sedk_mel_tables* tab, float* out, int64_t out_sb,
                               int64_t out_sm, int64_t out_st, int log_mode, float amin, float db_lo, float db_hi,
                               uint32_t* minmax, void* stream) {
    using namespace sedk;
    SEDK_PROF("logmel", (cudaStream_t)stream);
    SEDK_REQUIRE(wave && tab && out, "sedk_logmel_fwd: null pointer");
    SEDK_REQUIRE(B > 0, "sedk_logmel_fwd: B must be positive (got %d)", B);
    SEDK_REQUIRE(L > kHalf, "sedk_logmel_fwd: reflect padding needs L > %d samples (got %d)", kHalf, L);
    SEDK_REQUIRE(tab->hop > 0 && tab->hop <= kNfft, "sedk_logmel_fwd: hop %d out of range", tab->hop);
    SEDK_REQUIRE(tab->n_mels > 0 && tab->n_mels <= 256, "sedk_logmel_fwd: n_mels %d out of range", tab->n_mels);
    if (get_option("logmel_v2", 1) != 0)
        return launch_logmel_v2(wave, B, L, tab, out, out_sb, out_sm, out_st, log_mode, amin, db_lo, db_hi, minmax,
                                (cudaStream_t)stream);
    const int T = 1 + L / tab->hop;
    const int groups = cdiv(T, FR);
    SmemLayout lay = make_layout(tab->hop, tab->n_mels);
    SEDK_REQUIRE(lay.total <= 227 * 1024, "sedk_logmel_fwd: hop %d needs %zu B of shared memory", tab->hop, lay.total);
    int rc = opt_in_smem(logmel_kernel, lay.total);
    if (rc != SEDK_OK) return rc;
    long long total = (long long)B * groups;
    int per_sm = lay.total <= 110 * 1024 ? 2 : 1;
    int grid = (int)(total < (long long)num_sms() * per_sm ? total : (long long)num_sms() * per_sm);
    logmel_kernel<<<grid, NW * 32, lay.total, (cudaStream_t)stream>>>(wave, B, L, T, *tab, out, out_sb, out_sm, out_st,
                                                                    log_mode, amin, db_lo, db_hi, minmax, groups);
    SEDK_LAUNCH_CHECK("logmel_kernel");
    return SEDK_OK;
}
