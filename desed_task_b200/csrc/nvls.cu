// Data-parallel gradient all-reduce FUSED with the EMA + Adam update: one kernel over NVLink / NVSwitch peer memory, no NCCL
// call on the step (replaces ddp.allreduce_sum_ + sedk_adam_ema_dev at N > 1).
//
// The flat gradient of every rank lives in a symmetric allocation (same size on every rank, mapped into every peer's address
// space, plus - when the fabric supports it - one MULTICAST address that targets all copies at once).  One launch does
//
//   barrier A   every rank's backward has finished writing its local gradient
//   phase 1     reduce-scatter + all-gather: tile T belongs to rank T % world; the owner loads the SUM over all ranks
//               (multimem.ld_reduce: the switch adds the copies, NVLS) and stores it back to ALL copies (multimem.st).
//               Without a multicast address: the owner reads the peers' copies one by one (P2P, fixed rank order) and
//               writes the sum to each.  Either way every element is reduced exactly ONCE, so all replicas see the same
//               bits and stay identical (a one-shot "every rank reduces everything" would depend on the switch adding
//               in the same order for every requester).
//   barrier B   all copies hold the reduced gradient
//   phase 2     the fused EMA + Adam update of elementwise.cu on the local copy (bypassing L1: the data arrived from peers);
//               a rank's OWN tiles are updated already in phase 1, straight from the registers that hold the sum, so that
//               part of the update runs underneath the wait for the peers
//
// Barriers are per CTA INDEX: CTA c of a rank synchronises only with CTA c of the other ranks (flag[dst][slot][src] holds
// the launch epoch of src; see cross_barrier).  The tile -> CTA map makes that sufficient: in phase 2 CTA c touches exactly the tiles that the CTAs c of all ranks produced in phase 1.  CTAs of one
// GPU never wait for each other, so the launch cannot deadlock as long as every CTA is eventually scheduled (the grid is
// below one CTA per SM).
// do_adam = 0 stops after barrier B (plain all-reduce: the 2024 recipe clips by the global gradient norm before Adam).
#include "kernels.h"

namespace sedk {
namespace {

constexpr int NV_MAX_WORLD = 8;
constexpr int NV_GRID = 256;               // flag slots per barrier; the launch uses option "nvls_grid" (<= NV_GRID, default 128) CTAs
__device__ unsigned int nv_fault;           // set when a barrier wait exceeded NV_TIMEOUT_NS (a peer is gone): no endless spin
constexpr unsigned long long NV_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;
__device__ unsigned long long nv_dbg[8];   // globaltimer stamps of CTA 0 (option "nvls_debug"): start, barrier A, phase 1, barrier B, end
constexpr int NV_THREADS = 512;

struct NvArgs {
    float* gpeer[NV_MAX_WORLD];         // every rank's gradient copy in THIS rank's address space ([rank] = local)
    uint32_t* fpeer[NV_MAX_WORLD];      // every rank's flag block: [2 NV_GRID][NV_MAX_WORLD] slots + NV_GRID launch counters, zero at start
    float* gmc;                         // multicast address of the gradient copies, or NULL
    int rank, world;
};

// Barrier among the CTAs of one index: every rank STORES this launch's epoch into its slot of each peer's flag block
// (release, one-way: no round trip) and polls its own block until every peer's slot has reached the epoch (acquire).
// Epochs are per-CTA launch counters kept in the local block, so ranks agree as long as they make the same calls; a rank can
// run at most one barrier ahead of a peer, and the slots only ever grow (compared as a signed difference).  The first
// version used a CAS hand-shake (put 0 -> 1 on the peer, wait 1 -> 0 locally): two serialised NVLink round trips per barrier.
__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t val) {
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" :: "l"(addr), "r"(val) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void cross_barrier(const NvArgs& a, int slot, uint32_t epoch) {
    // bar.sync orders every thread's earlier stores (incl. multimem.st) before the signalling threads' release stores
    // (release is cumulative over the CTA barrier); a system fence in all 512 threads cost more than the barrier itself
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
        const int peer = threadIdx.x;
        st_release_sys(a.fpeer[peer] + (size_t)slot * NV_MAX_WORLD + a.rank, epoch);
        const uint32_t* local = a.fpeer[a.rank] + (size_t)slot * NV_MAX_WORLD + peer;
        unsigned long long t0 = 0;
        for (uint32_t spins = 1; (int32_t)(ld_acquire_sys(local) - epoch) < 0; spins++) {
            if ((spins & 0xfffu) == 0) {
                unsigned long long now;
                asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > NV_TIMEOUT_NS) { nv_fault = 1u; break; }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};\n"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_sys(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(float* p, const float4& v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};\n"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// EMA + Adam on one float4 of the flat buffers (elementwise.cu's arithmetic through common.cuh adam_elem / ema_elem)
struct NvHyper {
    float step_size, inv_sqrt_bc2, ema_alpha, grad_scale, beta1, beta2, eps;
};
__device__ __forceinline__ void update4(const float4& g4, float4& p4, float4& m4, float4& v4, float4& e4, bool has_ema,
                                        const NvHyper& h) {
    float* pp = reinterpret_cast<float*>(&p4);
    float* mm = reinterpret_cast<float*>(&m4);
    float* vv = reinterpret_cast<float*>(&v4);
    float* ee = reinterpret_cast<float*>(&e4);
    const float* gg = reinterpret_cast<const float*>(&g4);
    if (has_ema) {
#pragma unroll
        for (int k = 0; k < 4; k++) ee[k] = ema_elem(ee[k], pp[k], h.ema_alpha);
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
        adam_elem(pp[k], gg[k], mm[k], vv[k], h.step_size, h.beta1, h.beta2, h.eps, h.inv_sqrt_bc2, h.grad_scale);
}

// tile T (NV_THREADS float4): owner rank T % world, CTA (T / world) % gridDim.x
template <bool MC>
__global__ void __launch_bounds__(NV_THREADS)
allreduce_adam_kernel(NvArgs a, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                      float* __restrict__ ema, int64_t n4, int do_adam, float beta1, float beta2, float eps,
                      const float* __restrict__ hyper, int debug) {
    const int cta = blockIdx.x, tid = threadIdx.x, world = a.world;
    const int G = gridDim.x;
    auto stamp = [&](int k) {
        if (debug && cta == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
            nv_dbg[k] = t;
        }
    };
    stamp(0);
    const int64_t tiles = (n4 + NV_THREADS - 1) / NV_THREADS;
    __shared__ uint32_t s_epoch;
    if (tid == 0) {
        uint32_t* ctr = a.fpeer[a.rank] + (size_t)2 * NV_GRID * NV_MAX_WORLD + cta;     // this CTA's launch counter
        s_epoch = *ctr + 1u;
        *ctr = s_epoch;
    }
    __syncthreads();
    const uint32_t epoch = s_epoch;
    NvHyper hy = {0.f, 1.f, 0.f, 1.f, beta1, beta2, eps};
    if (do_adam) {
        hy.step_size = hyper[0]; hy.inv_sqrt_bc2 = hyper[1]; hy.ema_alpha = hyper[2]; hy.grad_scale = hyper[3];
    }
    cross_barrier(a, cta, epoch);
    stamp(1);
    // ---- phase 1: my tiles of this CTA index, four at a time (all loads of a batch in flight before the first store)
    constexpr int U = 4;
    const int64_t stride = (int64_t)G * world;
    for (int64_t T = (int64_t)cta * world + a.rank; T < tiles; T += U * stride) {
        float4 s[U];
        int64_t idx[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t i = (T + u * stride) * NV_THREADS + tid;
            idx[u] = (T + u * stride < tiles && i < n4) ? i : -1;
            if (idx[u] < 0) continue;
            if (MC) {
                s[u] = mc_ld_reduce(a.gmc + 4 * i);
            } else {
                s[u] = ld_sys(a.gpeer[0] + 4 * i);
                for (int r = 1; r < world; r++) {
                    const float4 t = ld_sys(a.gpeer[r] + 4 * i);
                    s[u].x += t.x; s[u].y += t.y; s[u].z += t.z; s[u].w += t.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (idx[u] < 0) continue;
            if (MC) {
                mc_st(a.gmc + 4 * idx[u], s[u]);
            } else {
                for (int r = 0; r < world; r++) st_sys(a.gpeer[r] + 4 * idx[u], s[u]);
            }
        }
        // the owner already holds the reduced gradient of its tiles: update them now, underneath the wait for the peers
        if (do_adam) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (idx[u] < 0) continue;
                const int64_t i = idx[u];
                float4 p4 = reinterpret_cast<const float4*>(p)[i], m4 = reinterpret_cast<const float4*>(m)[i];
                float4 v4 = reinterpret_cast<const float4*>(v)[i], e4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ema) e4 = reinterpret_cast<const float4*>(ema)[i];
                update4(s[u], p4, m4, v4, e4, ema != nullptr, hy);
                if (ema) reinterpret_cast<float4*>(ema)[i] = e4;
                reinterpret_cast<float4*>(p)[i] = p4;
                reinterpret_cast<float4*>(m)[i] = m4;
                reinterpret_cast<float4*>(v)[i] = v4;
            }
        }
    }
    stamp(2);
    cross_barrier(a, NV_GRID + cta, epoch);
    stamp(3);
    if (!do_adam) return;
    // ---- phase 2: the OTHER owners' tiles of this CTA index, on the local copy; two tiles per trip (ten 128-bit loads in
    //      flight per thread before the arithmetic)
    const float* g = a.gpeer[a.rank];
    for (int64_t T0 = (int64_t)cta * world; T0 < tiles; T0 += (int64_t)G * world) {
        for (int o = 0; o < world; o += 2) {
            float4 g4[2], p4[2], m4[2], v4[2], e4[2];
            int64_t idx[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int64_t i = (T0 + o + u) * NV_THREADS + tid;
                idx[u] = (o + u < world && o + u != a.rank && T0 + o + u < tiles && i < n4) ? i : -1;
                if (idx[u] < 0) continue;
                g4[u] = __ldcg(reinterpret_cast<const float4*>(g) + i);
                p4[u] = reinterpret_cast<const float4*>(p)[i];
                m4[u] = reinterpret_cast<const float4*>(m)[i];
                v4[u] = reinterpret_cast<const float4*>(v)[i];
                e4[u] = ema ? reinterpret_cast<const float4*>(ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (idx[u] < 0) continue;
                const int64_t i = idx[u];
                update4(g4[u], p4[u], m4[u], v4[u], e4[u], ema != nullptr, hy);
                if (ema) reinterpret_cast<float4*>(ema)[i] = e4[u];
                reinterpret_cast<float4*>(p)[i] = p4[u];
                reinterpret_cast<float4*>(m)[i] = m4[u];
                reinterpret_cast<float4*>(v)[i] = v4[u];
            }
        }
    }
    stamp(4);
}

}  // namespace
}  // namespace sedk

using namespace sedk;

extern "C" int64_t sedk_nvls_flag_bytes(void) { return (int64_t)(2 * NV_GRID * NV_MAX_WORLD + NV_GRID) * sizeof(uint32_t); }

extern "C" int sedk_allreduce_adam_nvls(float* p, float* m, float* v, float* ema, int64_t n, int do_adam, float beta1,
                                        float beta2, float eps, const float* hyper, void* g_mc, void* const* g_peers,
                                        void* const* flag_peers, int rank, int world, void* stream) {
    SEDK_PROF(do_adam ? "allreduce_adam_nvls" : "allreduce_nvls", (cudaStream_t)stream);
    SEDK_REQUIRE(world >= 1 && world <= NV_MAX_WORLD && rank >= 0 && rank < world, "nvls: bad rank / world (max %d ranks)",
                 NV_MAX_WORLD);
    SEDK_REQUIRE(g_peers && flag_peers && n > 0 && (n & 3) == 0, "nvls: n must be a positive multiple of 4 floats");
    SEDK_REQUIRE(!do_adam || (p && m && v && hyper), "nvls: the fused update needs p, m, v and the device scalars");
    NvArgs a;
    for (int r = 0; r < NV_MAX_WORLD; r++) {
        a.gpeer[r] = r < world ? (float*)g_peers[r] : nullptr;
        a.fpeer[r] = r < world ? (uint32_t*)flag_peers[r] : nullptr;
        SEDK_REQUIRE(r >= world || (a.gpeer[r] && a.fpeer[r] && ((uintptr_t)a.gpeer[r] & 15) == 0),
                     "nvls: peer %d pointers missing or not 16-byte aligned", r);
    }
    a.gmc = (float*)g_mc;
    SEDK_REQUIRE(((uintptr_t)g_mc & 15) == 0, "nvls: multicast address not 16-byte aligned");
    a.rank = rank;
    a.world = world;
    const int64_t n4 = n / 4;
    int grid = get_option("nvls_grid", 128);
    grid = grid < 1 ? 1 : (grid > NV_GRID ? NV_GRID : grid);
    const int debug = get_option("nvls_debug", 0);
    if (g_mc != nullptr)
        allreduce_adam_kernel<true><<<grid, NV_THREADS, 0, (cudaStream_t)stream>>>(a, p, m, v, ema, n4, do_adam, beta1, beta2,
                                                                                   eps, hyper, debug);
    else
        allreduce_adam_kernel<false><<<grid, NV_THREADS, 0, (cudaStream_t)stream>>>(a, p, m, v, ema, n4, do_adam, beta1, beta2,
                                                                                    eps, hyper, debug);
    SEDK_LAUNCH_CHECK("allreduce_adam_kernel");
    return SEDK_OK;
}

// diagnostic: the five globaltimer stamps (ns) CTA 0 took in the last launch with option "nvls_debug" = 1
extern "C" int sedk_nvls_debug_stamps(uint64_t* out5) {
    SEDK_REQUIRE(out5 != nullptr, "sedk_nvls_debug_stamps: null output");
    unsigned long long h[8];
    SEDK_CUDA(cudaMemcpyFromSymbol(h, nv_dbg, sizeof(h)));
    for (int i = 0; i < 5; i++) out5[i] = h[i];
    return SEDK_OK;
}

// 1 if a barrier of sedk_allreduce_adam_nvls timed out since the last call (then the buffers of this process are invalid); clears it
extern "C" int sedk_nvls_fault(void) {
    unsigned int h = 0, z = 0;
    if (cudaMemcpyFromSymbol(&h, nv_fault, sizeof(h)) != cudaSuccess) return -1;
    if (h != 0) cudaMemcpyToSymbol(nv_fault, &z, sizeof(z));
    return (int)h;
}
