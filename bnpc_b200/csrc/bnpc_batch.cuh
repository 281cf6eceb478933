// Chain-batched launches: every kernel of this library is a __device__ body plus ONE generic
// __global__ wrapper that runs the body for up to BATCH_MAX independent chains in one launch
// (blockIdx.z = chain, per-chain arguments and grid extents in the kernel parameter block).
//
// Why: the unit of parallelism of the reference is the chain (libs/MCMC.py:113-120 forks one
// process per chain).  On one GPU the chains step in lockstep; a step is a fixed sequence of
// small launches per chain, so issuing them chain by chain is bound by the launch rate, not by the
// GPU.  The composite entry points (bnpc_chain_*) therefore do not launch directly: a launch either
// goes out at once (no recorder active: the single-chain C ABI, batch of one) or is RECORDED into
// the calling thread's recorder under the current chain slot.  bnpc_batch_flush() then walks the
// recorded sequences of all chains in lockstep and merges launches of the same kernel into one
// batched launch.  Chains share no mutable buffer, so any interleaving that keeps each chain's own
// order is valid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <type_traits>
#include <utility>
#include <vector>

namespace bnpc {

constexpr int BATCH_MAX = 8;          // chains per batched launch (kernel parameter block <= 4 KB)
constexpr int GROUP_MAX = 64;         // chain slots of a recorder
constexpr int ARG_BYTES = 352;        // largest argument pack of a kernel body

// ---- argument packs ------------------------------------------------------------------------
template <typename... Ts> struct Pack;
template <> struct Pack<> {};
template <typename T, typename... Ts> struct Pack<T, Ts...> {
    T head;
    Pack<Ts...> tail;
};
template <size_t I, typename T, typename... Ts> struct PackGet {
    __host__ __device__ static const auto& get(const Pack<T, Ts...>& p) { return PackGet<I - 1, Ts...>::get(p.tail); }
};
template <typename T, typename... Ts> struct PackGet<0, T, Ts...> {
    __host__ __device__ static const T& get(const Pack<T, Ts...>& p) { return p.head; }
};
static inline void pack_fill(Pack<>&) {}
template <typename T, typename... Ts, typename A, typename... As>
static inline void pack_fill(Pack<T, Ts...>& p, const A& a, const As&... as) {
    p.head = (T)a;
    pack_fill(p.tail, as...);
}

template <typename F> struct FnTraits;
template <typename... Ts> struct FnTraits<void (*)(Ts...)> {
    using pack_t = Pack<std::decay_t<Ts>...>;
    template <typename... As> static pack_t make(const As&... as) {
        static_assert(sizeof...(As) == sizeof...(Ts), "argument count of a kernel launch");
        pack_t p;
        memset(&p, 0, sizeof(p));
        pack_fill(p, as...);
        return p;
    }
    template <void (*Body)(Ts...), size_t... I>
    __device__ __forceinline__ static void call(const pack_t& p, std::index_sequence<I...>) {
        Body(PackGet<I, std::decay_t<Ts>...>::get(p)...);
    }
    using seq = std::index_sequence_for<Ts...>;
};

template <int NB, typename P> struct Batch {
    unsigned gx[NB], gy[NB];
    P a[NB];
};

// ---- the wrappers -----------------------------------------------------------------------------
#define BNPC_WRAPPER_BODY                                                              \
    using T = FnTraits<decltype(Body)>;                                                \
    const int c = (NB == 1) ? 0 : (int)blockIdx.z;                                     \
    if (NB > 1) {                                                                      \
        if (blockIdx.x >= B.gx[c] || blockIdx.y >= B.gy[c]) return;                    \
    }                                                                                  \
    T::template call<Body>(B.a[c], typename T::seq{});

template <auto Body, int NB, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
batched_kernel(const __grid_constant__ Batch<NB, typename FnTraits<decltype(Body)>::pack_t> B) {
    BNPC_WRAPPER_BODY
}
// bodies without launch bounds
template <auto Body, int NB>
__global__ void batched_kernel_nb(const __grid_constant__ Batch<NB, typename FnTraits<decltype(Body)>::pack_t> B) {
    BNPC_WRAPPER_BODY
}

// ---- recorder -----------------------------------------------------------------------------------
struct Op {
    int kind;                     // 0 kernel, 1 memcpy, 2 event record
    const char* name;
    int (*merged)(Op* const* ops, int n, cudaStream_t s);   // identity + launcher of a kernel op
    unsigned gx, gy;
    int block;
    size_t smem;
    void* dst;
    const void* src;
    size_t bytes;
    cudaMemcpyKind mk;
    alignas(16) unsigned char args[ARG_BYTES];
};

struct Recorder {
    bool on = false;
    int cur = 0;
    std::vector<Op> q[GROUP_MAX];
};
static thread_local Recorder g_rec;

// per-kernel-name profile of the merged launches (events around every launch; a separate,
// untimed pass of bench.py)
struct ProfSlot {
    const char* name;
    cudaEvent_t a, b;
    int chains;
};
struct Profiler {
    bool on = false;
    std::vector<ProfSlot> slots;
    std::vector<cudaEvent_t> pool;
};
static thread_local Profiler g_prof;

static inline cudaEvent_t prof_event() {
    if (!g_prof.pool.empty()) {
        cudaEvent_t e = g_prof.pool.back();
        g_prof.pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

}  // namespace bnpc

// (defined in bnpc_kernels.cu)
static int fail(const char* what, cudaError_t e);
static int bad_arg(const char* what);
static std::atomic<long long> g_launches{0};

namespace bnpc {

template <typename K>
static int ensure_smem(K kernel, size_t smem, std::atomic<unsigned long long>& done, const char* name) {
    if (smem <= 48 * 1024) return 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return 0;
    // cudaFuncSetAttribute is per device; opt in to the largest dynamic shared memory of sm_100
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail(name, e);
    done.fetch_or(bit, std::memory_order_release);
    return 0;
}

template <auto Body, int NB, int MAXT, int MINB, typename P>
static int launch_wrapper(const char* name, const Batch<NB, P>& B, dim3 grid, int block, size_t smem,
                          cudaStream_t s, int chains) {
    static std::atomic<unsigned long long> done{0};
    cudaEvent_t ea = nullptr, eb = nullptr;
    if (g_prof.on) {
        ea = prof_event();
        eb = prof_event();
        cudaEventRecord(ea, s);
    }
    if constexpr (MAXT > 0) {
        if (int rc = ensure_smem(batched_kernel<Body, NB, MAXT, MINB>, smem, done, name)) return rc;
        batched_kernel<Body, NB, MAXT, MINB><<<grid, block, smem, s>>>(B);
    } else {
        if (int rc = ensure_smem(batched_kernel_nb<Body, NB>, smem, done, name)) return rc;
        batched_kernel_nb<Body, NB><<<grid, block, smem, s>>>(B);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(name, e);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (g_prof.on) {
        cudaEventRecord(eb, s);
        g_prof.slots.push_back(ProfSlot{name, ea, eb, chains});
    }
    static const int sync_each = getenv("BNPC_SYNC_EACH") ? 1 : 0;      // debugging
    if (sync_each) cudaStreamSynchronize(s);
    return 0;
}

// one batched launch for the same kernel op of n <= BATCH_MAX chains
template <auto Body, int MAXT, int MINB>
static int launch_merged(Op* const* ops, int n, cudaStream_t s) {
    using P = typename FnTraits<decltype(Body)>::pack_t;
    static const int force_nb8 = getenv("BNPC_FORCE_NB8") ? 1 : 0;      // debugging
    if (n == 1 && !force_nb8) {
        Batch<1, P> B;
        B.gx[0] = ops[0]->gx; B.gy[0] = ops[0]->gy;
        memcpy(&B.a[0], ops[0]->args, sizeof(P));
        return launch_wrapper<Body, 1, MAXT, MINB>(ops[0]->name, B, dim3(ops[0]->gx, ops[0]->gy, 1), ops[0]->block,
                                                   ops[0]->smem, s, 1);
    }
    Batch<BATCH_MAX, P> B;
    memset(&B, 0, sizeof(B));
    unsigned gx = 1, gy = 1;
    size_t smem = 0;
    for (int i = 0; i < n; ++i) {
        B.gx[i] = ops[i]->gx; B.gy[i] = ops[i]->gy;
        memcpy(&B.a[i], ops[i]->args, sizeof(P));
        gx = ops[i]->gx > gx ? ops[i]->gx : gx;
        gy = ops[i]->gy > gy ? ops[i]->gy : gy;
        smem = ops[i]->smem > smem ? ops[i]->smem : smem;
    }
    return launch_wrapper<Body, BATCH_MAX, MAXT, MINB>(ops[0]->name, B, dim3(gx, gy, n), ops[0]->block, smem, s, n);
}

template <auto Body, int MAXT, int MINB, typename... As>
static int launch(const char* name, dim3 grid, int block, size_t smem, cudaStream_t s, const As&... as) {
    using T = FnTraits<decltype(Body)>;
    using P = typename T::pack_t;
    static_assert(sizeof(P) <= ARG_BYTES, "argument pack larger than ARG_BYTES");
    static_assert(std::is_trivially_copyable<P>::value, "kernel arguments must be plain data");
    static_assert(sizeof(Batch<BATCH_MAX, P>) <= 4000, "parameter block of a batched launch exceeds 4 KB");
    if (grid.z != 1) return bad_arg("kernel bodies use blockIdx.x / .y only (z is the chain)");
    if (grid.x == 0 || grid.y == 0) return 0;
    if (g_rec.on) {
        g_rec.q[g_rec.cur].emplace_back();
        Op& op = g_rec.q[g_rec.cur].back();
        op.kind = 0; op.name = name; op.merged = &launch_merged<Body, MAXT, MINB>;
        op.gx = grid.x; op.gy = grid.y; op.block = block; op.smem = smem;
        const P p = T::make(as...);
        memcpy(op.args, &p, sizeof(P));
        return 0;
    }
    Batch<1, P> B;
    B.gx[0] = grid.x; B.gy[0] = grid.y;
    B.a[0] = T::make(as...);
    return launch_wrapper<Body, 1, MAXT, MINB>(name, B, grid, block, smem, s, 1);
}

}  // namespace bnpc

#define BNPC_LAUNCH(body, maxt, minb, grid, block, smem, stream, ...)                                       \
    do {                                                                                                     \
        if (int rc__ = bnpc::launch<body, maxt, minb>(#body, dim3(grid), (int)(block), (size_t)(smem),       \
                                                      (cudaStream_t)(stream), __VA_ARGS__))                  \
            return rc__;                                                                                     \
    } while (0)
