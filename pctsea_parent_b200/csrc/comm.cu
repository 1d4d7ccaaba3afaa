// comm.cu — the one exchange step of the path, inside the C ABI: the per-rank null slices are all-gathered with
// NCCL (NVLink / NVSwitch) on the ctx stream, so a Java host gets the multi-GPU path without any Python plumbing.
//
// Replaces nothing in the reference (it is single-process: the permutation loop PCTSEA.java:1398-1429 runs all P
// permutations on one JVM); it realises SURVEY.md §8(e): permutations shard by global index, the null slices
// [T][P_r] are the only data exchanged, a10-a12 then run on the complete null on every rank.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy the host process already loaded, e.g. PyTorch's, or
// the system one), so single-GPU users of libpctsea_b200.so do not need NCCL at all. Only entry points whose
// signatures are stable across NCCL 2.x are used.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>   // types only; no link-time dependency

namespace pctsea {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string load_error;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            const char* e = dlerror();
            api.load_error = std::string("cannot load libnccl.so.2: ") + (e ? e : "?");
            return;
        }
        auto sym = [&](const char* s) -> void* {
            void* p = dlsym(api.handle, s);
            if (!p && api.load_error.empty()) api.load_error = std::string("libnccl lacks ") + s;
            return p;
        };
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

static NcclApi& nccl_or_throw() {
    NcclApi& a = nccl_api();
    if (!a.load_error.empty()) throw Error(PCTSEA_ENCCL, a.load_error);
    return a;
}

#define PCT_NCCL(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            char _b[512];                                                                           \
            snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #expr, nccl_api().GetErrorString(_r),  \
                     __FILE__, __LINE__);                                                           \
            throw ::pctsea::Error(PCTSEA_ENCCL, _b);                                                \
        }                                                                                           \
    } while (0)

static_assert(sizeof(ncclUniqueId) == PCTSEA_COMM_ID_BYTES, "ncclUniqueId is 128 bytes in every NCCL 2.x");

void comm_unique_id(void* id_out) {
    NcclApi& a = nccl_or_throw();
    ncclUniqueId id;
    PCT_NCCL(a.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
}

void comm_init_rank(pctsea_ctx* c, int32_t nranks, int32_t rank, const void* id_bytes) {
    NcclApi& a = nccl_or_throw();
    PCT_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "rank / nranks out of range");
    PCT_REQUIRE(id_bytes != nullptr, "unique id is NULL");
    PCT_REQUIRE(c->comm == nullptr, "this ctx already has a communicator");
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    ncclComm_t comm = nullptr;
    PCT_NCCL(a.CommInitRank(&comm, nranks, id, rank));   // collective over the ranks; binds to the ctx's device
    c->comm = comm;
    c->comm_nranks = nranks;
    c->comm_rank = rank;
}

void comm_destroy(pctsea_ctx* c) {
    if (!c->comm) return;
    cudaStreamSynchronize(c->stream);
    nccl_api().CommDestroy(reinterpret_cast<ncclComm_t>(c->comm));
    c->comm = nullptr;
    c->comm_nranks = 1;
    c->comm_rank = 0;
}

// (begin, count) of a rank's contiguous share of n items: sizes differ by at most one (the same split as
// pctsea_parent_b200/dist.py perm_slice)
void comm_share(int64_t n, int32_t nranks, int32_t rank, int64_t* begin, int64_t* count) {
    const int64_t base = n / nranks, rem = n % nranks;
    *begin = rank * base + std::min<int64_t>(rank, rem);
    *count = base + (rank < rem ? 1 : 0);
}

// In-place all-gather of `bytes_per_rank` bytes per rank on the ctx stream: rank r's block sits at buf + r * bytes_per_rank.
void comm_all_gather_in_place(pctsea_ctx* c, void* buf, size_t bytes_per_rank) {
    NcclApi& a = nccl_or_throw();
    PCT_REQUIRE(c->comm != nullptr, "no communicator");
    if (bytes_per_rank == 0) return;
    PCT_NCCL(a.AllGather(static_cast<char*>(buf) + (size_t)c->comm_rank * bytes_per_rank, buf, bytes_per_rank, ncclChar,
                         reinterpret_cast<ncclComm_t>(c->comm), c->stream));
}

// gathered[r][t][width] (rank r's slice in its first count_r columns) -> pooled[t][P], permutations in global order
__global__ void k_pool_null(const float* __restrict__ gathered, int32_t T, int32_t P, int32_t width, int32_t nranks,
                            float* __restrict__ pooled) {
    const int64_t n = (int64_t)T * P;
    const int32_t base = P / nranks, rem = P % nranks;
    for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < n; id += (int64_t)gridDim.x * blockDim.x) {
        const int32_t t = (int32_t)(id / P), p = (int32_t)(id - (int64_t)t * P);
        // the first `rem` ranks own base + 1 permutations each
        const int32_t cut = rem * (base + 1);
        int32_t r, off;
        if (p < cut) { r = p / (base + 1); off = p - r * (base + 1); }
        else { r = rem + (p - cut) / max(base, 1); off = (p - cut) - (r - rem) * base; }
        pooled[id] = gathered[((size_t)r * T + t) * width + off];
    }
}

// All-gather of the local null slice(s) [T][width] on the ctx stream, then the reorder into [T][P].
void comm_pool_null(pctsea_ctx* c, int32_t T, int32_t P, int32_t width, bool with_dstat) {
    NcclApi& a = nccl_or_throw();
    ncclComm_t comm = reinterpret_cast<ncclComm_t>(c->comm);
    const size_t slice = (size_t)T * width;
    const int R = c->comm_nranks;
    c->null_gather.reserve(slice * R * 4 * (with_dstat ? 2 : 1) + 64);
    c->null_pool.reserve((size_t)std::max(T, 1) * std::max(P, 1) * 4 + 64);
    if (with_dstat) c->null_pool_d.reserve((size_t)std::max(T, 1) * std::max(P, 1) * 4 + 64);
    if (slice == 0) return;
    float* g0 = c->null_gather.as<float>();
    float* g1 = g0 + slice * R;
    if (with_dstat) PCT_NCCL(a.GroupStart());
    PCT_NCCL(a.AllGather(c->null_ews.p, g0, slice, ncclFloat32, comm, c->stream));
    if (with_dstat) {
        PCT_NCCL(a.AllGather(c->null_d.p, g1, slice, ncclFloat32, comm, c->stream));
        PCT_NCCL(a.GroupEnd());
    }
    const int64_t n = (int64_t)T * P;
    const int grid = (int)std::min<int64_t>(div_up(n, 256), (int64_t)c->num_sms * 8);
    k_pool_null<<<grid, 256, 0, c->stream>>>(g0, T, P, width, R, c->null_pool.as<float>());
    if (with_dstat) k_pool_null<<<grid, 256, 0, c->stream>>>(g1, T, P, width, R, c->null_pool_d.as<float>());
    count_launch(c, with_dstat ? 2 : 1);
    PCT_CUDA(cudaGetLastError());
}

}  // namespace pctsea
