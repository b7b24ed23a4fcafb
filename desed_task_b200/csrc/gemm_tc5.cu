// Dense GEMMs of the GRU path on tcgen05 / TMEM / TMA (sm_100a, TF32 multiplies, fp32 accumulate):
//   * input projections   gi_d = x W_ih,d^T + b_ih,d        (RNN.py:19-30; "NT": both operands K-major; the two
//                                                             directions are two problems of one launch, blockIdx.z)
//   * input gradients     dx = dgi_f W_ih,f + dgi_b W_ih,b   ("NN": B = W_ih [K, N] is MN-major; the two directions are
//                                                             ONE problem: split-K over two operand pairs into one
//                                                             TMEM accumulator, so dx is written once, no beta pass)
// One CTA = one 128 x 128 output tile; K streamed in 32-wide chunks through a 4-stage TMA / mbarrier ring; warp 0 = TMA
// producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue (tcgen05.ld -> + bias -> 128-bit stores).
// Operand layouts: K-major tiles are [128 rows x 32 k] with the 128-byte swizzle; the MN-major B tile is 4 chunks of
// [32 k-rows x 32 n] with the 32-byte-atom 128-byte swizzle TF32 MN-major operands require (same as conv_wgrad_tc5_kernel).
#include "kernels.h"
#include "tc5.cuh"

namespace sedk {
namespace {

constexpr int GM_STAGES = 4;
constexpr int GM_TILE_BYTES = 128 * 128;                  // one operand tile of one stage: 16 KB
constexpr int GM_STAGE_BYTES = 2 * GM_TILE_BYTES;
constexpr int GM_THREADS = 192;
constexpr size_t GM_SMEM = (size_t)GM_STAGES * GM_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

// MODE 0: blockIdx.z selects one of two independent problems (A0,B0,C0,bias0) / (A1,B1,C1,bias1)
// MODE 1: one problem, C0 = A0 B0 + A1 B1 (split-K over the two operand pairs)
template <bool BMN, int MODE>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_tc5_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1, float* __restrict__ C0,
                float* __restrict__ C1, const float* __restrict__ bias0, const float* __restrict__ bias1, int M, int N, int K,
                int ldc) {
    pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* aligned = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(aligned + (size_t)GM_STAGES * GM_STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + GM_STAGES;
    uint64_t* accum = bars + 2 * GM_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GM_STAGES + 1);
    constexpr uint32_t IDESC = tc_idesc(128) | (BMN ? (1u << 16) : 0u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;
    const int z = MODE == 0 ? (int)blockIdx.z : 0;
    const int kchunks = K / 32;
    const int nit = MODE == 1 ? 2 * kchunks : kchunks;

    if (tid == 0) {
        for (int s = 0; s < GM_STAGES; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(accum, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                     "r"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc5_fence_before();
    __syncthreads();
    tc5_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < nit; it++) {
                const int s = it % GM_STAGES, ph = (it / GM_STAGES) & 1;
                const int pair = MODE == 1 ? (it >= kchunks ? 1 : 0) : z;
                const int kc = MODE == 1 && it >= kchunks ? it - kchunks : it;
                const CUtensorMap* ta = pair ? &tmA1 : &tmA0;
                const CUtensorMap* tb = pair ? &tmB1 : &tmB0;
                mbar_wait_u32(smem_u32(&empty[s]), ph ^ 1);
                const uint32_t a_dst = base + s * GM_STAGE_BYTES, b_dst = a_dst + GM_TILE_BYTES;
                mbar_expect_tx(&full[s], GM_STAGE_BYTES);
                tma_load_2d(a_dst, ta, smem_u32(&full[s]), kc * 32, m0);
                if (BMN) {
#pragma unroll
                    for (int c = 0; c < 4; c++) tma_load_2d(b_dst + c * 4096, tb, smem_u32(&full[s]), n0 + 32 * c, kc * 32);
                } else {
                    tma_load_2d(b_dst, tb, smem_u32(&full[s]), kc * 32, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int it = 0; it < nit; it++) {
                const int s = it % GM_STAGES, ph = (it / GM_STAGES) & 1;
                mbar_wait_u32(smem_u32(&full[s]), ph);
                tc5_fence_after();
                const uint32_t a_src = base + s * GM_STAGE_BYTES, b_src = a_src + GM_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t da = umma_desc_sw128(a_src + k * 32);
                    const uint64_t db = BMN ? umma_desc_mn_sw128(b_src + k * 1024, 4096) : umma_desc_sw128(b_src + k * 32);
                    umma_tf32(tmem, da, db, IDESC, (it | k) != 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(&empty[s]));
            }
            umma_commit(smem_u32(accum));
        }
    } else {
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        float* C = z ? C1 : C0;
        const float* bias = z ? bias1 : bias0;
        mbar_wait_u32(smem_u32(accum), 0);
        tc5_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
            const int n = n0 + c * 32;
            if (n >= N) break;
            uint32_t v[32];
            tmem_ld32(v, tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32));
            if (m < M) {
                float* crow = C + (size_t)m * ldc + n;
                if (n + 32 <= N) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        if (bias != nullptr) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n) + j);
                            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                        }
                        reinterpret_cast<float4*>(crow)[j] = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (n + j < N) crow[j] = __uint_as_float(v[j]) + (bias != nullptr ? bias[n + j] : 0.f);
                }
            }
        }
        tc5_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc5_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(128) : "memory");
    }
}

// row-major [rows, cols] fp32 matrix, box {32 cols, box_rows}
int map_2d(CUtensorMap* tm, const float* p, int rows, int cols, int ld, int box_rows, bool atom32) {
    EncodeTiledFn enc = encode_fn();
    SEDK_REQUIRE(enc != nullptr, "gemm_tc5: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SEDK_REQUIRE(r == CUDA_SUCCESS, "gemm_tc5: cuTensorMapEncodeTiled failed with %d", (int)r);
    return SEDK_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <class Kern>
int opt_in_once(Kern k, bool& done) {
    if (!done) {
        int rc = opt_in_smem(k, GM_SMEM);
        if (rc) return rc;
        done = true;
    }
    return SEDK_OK;
}

}  // namespace

bool gemm_tc5_ok(int M, int N, int K, int precision) {
    return precision == 0 && tc5_enabled() && get_option("gemm_tc5", 1) != 0 && M >= 128 && K >= 32 && K % 32 == 0 &&
           N % 4 == 0 && N >= 32;
}

// C_d[M, N] = A[M, K] B_d[N, K]^T + bias_d   for d = 0, 1 (A shared; lda = K, ldb = K, ldc = N)
int launch_gemm_tc5_nt2(const float* A, const float* const B[2], const float* const bias[2], float* const C[2], int M, int N,
                        int K, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "gemm_tc5_NT_%dx%dx%d_x2", M, N, K);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(aligned16(A) && aligned16(B[0]) && aligned16(B[1]) && aligned16(C[0]) && aligned16(C[1]) &&
                     aligned16(bias[0]) && aligned16(bias[1]),
                 "gemm_tc5: operands must be 16-byte aligned");
    CUtensorMap ta, tb0, tb1;
    int rc = map_2d(&ta, A, M, K, K, 128, false);
    if (rc) return rc;
    rc = map_2d(&tb0, B[0], N, K, K, 128, false);
    if (rc) return rc;
    rc = map_2d(&tb1, B[1], N, K, K, 128, false);
    if (rc) return rc;
    static bool cfg = false;
    auto kern = gemm_tc5_kernel<false, 0>;
    rc = opt_in_once(kern, cfg);
    if (rc) return rc;
    dim3 grid(cdiv(M, 128), cdiv(N, 128), 2);
    SEDK_CUDA(pdl_launch(kern, dim3(grid), dim3(GM_THREADS), (size_t)(GM_SMEM), s, ta, tb0, ta, tb1, C[0], C[1], bias[0], bias[1], M, N, K, N));
    SEDK_LAUNCH_CHECK("gemm_tc5_kernel<NT>");
    return SEDK_OK;
}

// C[M, N] = A[M, K] B[N, K]^T + bias   (one problem of the kernel above; lda = K, ldb = K, ldc = N)
int launch_gemm_tc5_nt1(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "gemm_tc5_NT_%dx%dx%d_x1", M, N, K);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(aligned16(A) && aligned16(B) && aligned16(C) && (bias == nullptr || aligned16(bias)),
                 "gemm_tc5: operands must be 16-byte aligned");
    CUtensorMap ta, tb;
    int rc = map_2d(&ta, A, M, K, K, 128, false);
    if (rc) return rc;
    rc = map_2d(&tb, B, N, K, K, 128, false);
    if (rc) return rc;
    static bool cfg = false;
    auto kern = gemm_tc5_kernel<false, 0>;
    rc = opt_in_once(kern, cfg);
    if (rc) return rc;
    dim3 grid(cdiv(M, 128), cdiv(N, 128), 1);
    SEDK_CUDA(pdl_launch(kern, dim3(grid), dim3(GM_THREADS), (size_t)(GM_SMEM), s, ta, tb, ta, tb, C, C, bias, bias, M, N, K, N));
    SEDK_LAUNCH_CHECK("gemm_tc5_kernel<NT1>");
    return SEDK_OK;
}

// C[M, N] = A0[M, K] B0[K, N] + A1[M, K] B1[K, N]   (lda = K, ldb = N, ldc = N)
int launch_gemm_tc5_nn_pair(const float* const A[2], const float* const B[2], float* C, int M, int N, int K, cudaStream_t s) {
    char pname[64];
    snprintf(pname, sizeof(pname), "gemm_tc5_NN_%dx%dx%d_k2", M, N, K);
    SEDK_PROF(pname, s);
    SEDK_REQUIRE(aligned16(A[0]) && aligned16(A[1]) && aligned16(B[0]) && aligned16(B[1]) && aligned16(C),
                 "gemm_tc5: operands must be 16-byte aligned");
    CUtensorMap ta0, ta1, tb0, tb1;
    int rc = map_2d(&ta0, A[0], M, K, K, 128, false);
    if (rc) return rc;
    rc = map_2d(&ta1, A[1], M, K, K, 128, false);
    if (rc) return rc;
    rc = map_2d(&tb0, B[0], K, N, N, 32, true);
    if (rc) return rc;
    rc = map_2d(&tb1, B[1], K, N, N, 32, true);
    if (rc) return rc;
    static bool cfg = false;
    auto kern = gemm_tc5_kernel<true, 1>;
    rc = opt_in_once(kern, cfg);
    if (rc) return rc;
    dim3 grid(cdiv(M, 128), cdiv(N, 128), 1);
    SEDK_CUDA(pdl_launch(kern, dim3(grid), dim3(GM_THREADS), (size_t)(GM_SMEM), s, ta0, tb0, ta1, tb1, C, C, nullptr, nullptr, M, N, K, N));
    SEDK_LAUNCH_CHECK("gemm_tc5_kernel<NN>");
    return SEDK_OK;
}

}  // namespace sedk
