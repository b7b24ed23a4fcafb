// Shared device/host helpers for libsedk (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/sedk.h"

namespace sedk {

// ---------------------------------------------------------------------------------------------
// error plumbing (no exceptions cross the C ABI)
void set_error(const char* fmt, ...);
int  check_launch(const char* what);
void count_launch();
bool profiling_on();
// named kernel-variant switches (sedk_set_option / env SEDK_<NAME>)
int  get_option(const char* name, int dflt);
void set_option(const char* name, int value);
// RAII device-timing scope around one launcher (no-op unless sedk_profile_enable(1) and the stream is not capturing)
struct ProfScope {
    ProfScope(const char* name, cudaStream_t s);
    ~ProfScope();
    int idx_;
    cudaStream_t s_;
};
#define SEDK_PROF(name, stream) ::sedk::ProfScope prof_scope__(name, stream)

#define SEDK_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            ::sedk::set_error(__VA_ARGS__);                       \
            return SEDK_ERR_INVALID;                              \
        }                                                         \
    } while (0)

#define SEDK_UNSUPPORTED(...)                                     \
    do {                                                          \
        ::sedk::set_error(__VA_ARGS__);                           \
        return SEDK_ERR_UNSUPPORTED;                              \
    } while (0)

#define SEDK_CUDA(call)                                                          \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) {                                                \
            ::sedk::set_error("%s failed: %s", #call, cudaGetErrorString(e__));  \
            return SEDK_ERR_CUDA;                                                \
        }                                                                        \
    } while (0)

#define SEDK_LAUNCH_CHECK(what)                                   \
    do {                                                          \
        int rc__ = ::sedk::check_launch(what);                    \
        if (rc__ != SEDK_OK) return rc__;                         \
    } while (0)

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <class K>
inline int opt_in_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", bytes, cudaGetErrorString(e));
            return SEDK_ERR_CUDA;
        }
    }
    return SEDK_OK;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
// launch on the dependency chain with programmatic stream serialisation (option "pdl"; default 0 = plain launch order:
// measured on B200, the early-scheduled dependents cost more than they hide - supervised step 2.205 ms with vs 2.188 ms
// without, mean-teacher 4.469 vs 4.384 ms, where the waiting CTAs take SM slots from the parallel graph branches).
// ONLY for kernels whose first statement is pdl_enter().
template <class... KArgs, class... Args>
inline cudaError_t pdl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    const int on = get_option("pdl", 0) != 0 ? 1 : 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = on;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---------------------------------------------------------------------------------------------
// order-preserving float <-> uint32 map (for atomicMin / atomicMax on floats)
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  The kernels on the step's dependency chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (pdl_launch below): such a kernel may be scheduled while its
// predecessor in the stream is still draining, so its FIRST statement must be pdl_enter(): it lets ITS successor start
// launching (launch_dependents) and then blocks until every prerequisite grid has completed and flushed its memory
// (griddepcontrol.wait).  Nothing before that line may touch global memory.  The launch latency / CTA scheduling of
// kernel k+1 thereby overlaps the tail of kernel k (~80 dependent launches per training step).
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
// MUFU-based variant (ex2 + rcp): abs error ~1e-7, used inside the fused CNN block kernels and the recurrence
// (raw ex2.approx / rcp.approx: none of the range-check code __expf / __fdividef carry; ex2 -> inf gives rcp -> 0, the limit)
__device__ __forceinline__ float fast_sigmoidf_(float x) {
    float e, y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(1.0f + e));
    return y;
}

// ---------------------------------------------------------------------------------------------
// cp.async (LDGSTS) 16-byte copies with zero fill
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    int src_bytes = valid ? 16 : 0;   // src-size 0 -> destination zero-filled, source not read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// 1-D bulk async copy (TMA engine, SASS UBLKCP) + mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes));
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// TF32 tensor-core MMA (mma.sync m16n8k8), optional 3xTF32 error compensation.
// Fragment layout (g = lane>>2, t = lane&3):
//   A: a0=(row g, col t) a1=(g+8, t) a2=(g, t+4) a3=(g+8, t+4);  B: b0=(k t, n g) b1=(k t+4, n g)
//   C: c0=(g, 2t) c1=(g, 2t+1) c2=(g+8, 2t) c3=(g+8, 2t+1)
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return r;
}
// Operand rounding for mma.sync TF32 without the three-instruction sequence cvt.rna.tf32 compiles to on sm_100a (FSETP against
// inf + conditional add + mask): the MMA unit ignores the 13 low mantissa bits of an fp32 operand, so adding half a TF32
// ulp to the magnitude IS round-to-nearest-ties-away for every finite value (inf / nan do not occur in these operands).
// ONLY for values that go straight into mma_tf32; anything stored for another consumer keeps to_tf32 (clean low bits).
__device__ __forceinline__ uint32_t to_tf32_mma(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One k8 step of a warp tile: acc[MF][NF] += A(MF x 16 rows, 8 k) * B(8 k, NF x 8 cols).
// fa(i, r, c): value of A at m-fragment i, row r in {0 (=g), 1 (=g+8)}, k-col c in {0 (=t), 1 (=t+4)}
// fb(j, c)   : value of B at n-fragment j, k-row c in {0 (=t), 1 (=t+4)}, column g
template <int MF, int NF, bool X3, class FA, class FB>
__device__ __forceinline__ void warp_mma_k8(float (&acc)[MF][NF][4], FA fa, FB fb) {
    uint32_t ah[MF][4], al[MF][4];
#pragma unroll
    for (int i = 0; i < MF; i++) {
        float v0 = fa(i, 0, 0), v1 = fa(i, 1, 0), v2 = fa(i, 0, 1), v3 = fa(i, 1, 1);
        ah[i][0] = to_tf32(v0); ah[i][1] = to_tf32(v1); ah[i][2] = to_tf32(v2); ah[i][3] = to_tf32(v3);
        if (X3) {
            al[i][0] = to_tf32(v0 - __uint_as_float(ah[i][0]));
            al[i][1] = to_tf32(v1 - __uint_as_float(ah[i][1]));
            al[i][2] = to_tf32(v2 - __uint_as_float(ah[i][2]));
            al[i][3] = to_tf32(v3 - __uint_as_float(ah[i][3]));
        }
    }
#pragma unroll
    for (int j = 0; j < NF; j++) {
        float w0 = fb(j, 0), w1 = fb(j, 1);
        uint32_t bh[2] = {to_tf32(w0), to_tf32(w1)};
        uint32_t bl[2];
        if (X3) {
            bl[0] = to_tf32(w0 - __uint_as_float(bh[0]));
            bl[1] = to_tf32(w1 - __uint_as_float(bh[1]));
        }
#pragma unroll
        for (int i = 0; i < MF; i++) {
            if (X3) {
                mma_tf32(acc[i][j], al[i], bh);
                mma_tf32(acc[i][j], ah[i], bl);
            }
            mma_tf32(acc[i][j], ah[i], bh);
        }
    }
}

// Same step with the fragments fetched by ldmatrix (one LDSM.x4 per A fragment / per pair of B fragments instead of
// 4 + 4 scalar LDS).  Valid when BOTH operands are K-major in shared memory with 16-byte aligned rows: a 16x8 TF32 A
// fragment is four 8x8 b16 matrices whose 32-bit words are exactly (row g, k t), (g+8, t), (g, t+4), (g+8, t+4).
// a_addr(i): shared byte address THIS lane contributes for m-fragment i, i.e. of row (lane&7) + 8*((lane>>3)&1), k-offset
// 4*(lane>>4);  b_addr(jp): for the fragment pair (2jp, 2jp+1): row n = (2jp + (lane>>4))*8 + (lane&7), k-offset
// 4*((lane>>3)&1).
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
template <int MF, int NF, bool X3, class FA, class FB>
__device__ __forceinline__ void warp_mma_k8_ldsm(float (&acc)[MF][NF][4], FA a_addr, FB b_addr) {
    static_assert(NF % 2 == 0, "fragment pairs");
    uint32_t ah[MF][4], al[MF][4];
#pragma unroll
    for (int i = 0; i < MF; i++) {
        uint32_t raw[4];
        ldsm_x4(raw, a_addr(i));
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float v = __uint_as_float(raw[q]);
            ah[i][q] = to_tf32(v);
            if (X3) al[i][q] = to_tf32(v - __uint_as_float(ah[i][q]));
        }
    }
#pragma unroll
    for (int jp = 0; jp < NF / 2; jp++) {
        uint32_t raw[4];
        ldsm_x4(raw, b_addr(jp));
        uint32_t bh[2][2], bl[2][2];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float v = __uint_as_float(raw[q]);
            bh[q >> 1][q & 1] = to_tf32(v);
            if (X3) bl[q >> 1][q & 1] = to_tf32(v - __uint_as_float(bh[q >> 1][q & 1]));
        }
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int i = 0; i < MF; i++) {
                if (X3) {
                    mma_tf32(acc[i][2 * jp + h], al[i], bh[h]);
                    mma_tf32(acc[i][2 * jp + h], ah[i], bl[h]);
                }
                mma_tf32(acc[i][2 * jp + h], ah[i], bh[h]);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-7 counter RNG (the 7-round variant of Salmon et al., which still passes BigCrush; dropout masks are
// regenerated in backward from (seed, stream, index) instead of being stored)
struct Philox {
    uint32_t key0, key1;
    __device__ __forceinline__ Philox(uint64_t seed) : key0((uint32_t)seed), key1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint64_t index, uint64_t stream) const {
        uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
        uint32_t k0 = key0, k1 = key1;
#pragma unroll
        for (int r = 0; r < 7; r++) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
// keep-probability test for one 32-bit draw: keep iff u >= p  (u uniform in [0,1))
__device__ __forceinline__ bool keep_from_bits(uint32_t bits, uint32_t thresh) { return bits >= thresh; }
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
    double t = (double)p * 4294967296.0;
    if (t <= 0.0) return 0u;
    if (t >= 4294967295.0) return 0xffffffffu;
    return (uint32_t)t;
}
// keep decision for element `e` (any 64-bit linear index) of dropout stream `stream`
__device__ __forceinline__ bool dropout_keep(const Philox& ph, uint64_t e, uint64_t stream, uint32_t thresh) {
    uint4 r = ph(e >> 2, stream);
    uint32_t sel = (uint32_t)(e & 3);
    uint32_t bits = sel == 0 ? r.x : sel == 1 ? r.y : sel == 2 ? r.z : r.w;
    return bits >= thresh;
}
// ---------------------------------------------------------------------------------------------
// One element of the fused EMA + Adam update (include/sedk.h: sedk_adam_ema).  Written with explicit roundings so that
// every kernel that applies it (elementwise.cu adam_ema_kernel, nvls.cu allreduce_adam_kernel) produces the same bits -
// left to the compiler, FMA contraction differed between the two.
__device__ __forceinline__ float ema_elem(float ema, float p, float ema_alpha) {
    return __fmaf_rn(ema, ema_alpha, __fmul_rn(p, 1.0f - ema_alpha));
}
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float step_size, float beta1, float beta2,
                                          float eps, float inv_sqrt_bc2, float grad_scale) {
    const float gi = __fmul_rn(g, grad_scale);
    const float mi = __fmaf_rn(__fsub_rn(gi, m), 1.0f - beta1, m);                  // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = __fmaf_rn(v, beta2, __fmul_rn(__fmul_rn(1.0f - beta2, gi), gi));
    m = mi;
    v = vi;
    const float denom = __fmaf_rn(sqrtf(vi), inv_sqrt_bc2, eps);
    p = __fmaf_rn(-step_size, __fdiv_rn(mi, denom), p);
}
#endif  // __CUDACC__

}  // namespace sedk
