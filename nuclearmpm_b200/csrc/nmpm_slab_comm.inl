// Native slab step: the whole per-step protocol of a multi-GPU x-slab (see include/nmpm.h and
// nuclearmpm_b200/slab.py, which is the reference implementation of the same protocol over
// torch.distributed) driven from C++ with NCCL point-to-point calls on the sim's stream.
//
// Included at the end of nmpm_api.cu (it needs nmpm_sim and the static step helpers).  NCCL is not a
// link-time dependency: the symbols are taken with dlopen/dlsym from the libnccl.so.2 the host process
// already has loaded (torch's), or from an explicit path.
#include <cstdlib>
#include <dlfcn.h>
#if __has_include(<nccl.h>)
#include <nccl.h>
#else
// NCCL is only ever reached through dlopen at run time; without its development header the handful of types
// the entry points below use are declared here (ABI of NCCL 2.x), so the single-GPU library still builds.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt32 = 2, ncclFloat32 = 7 } ncclDataType_t;
#endif

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int nccl_load(const char* path, std::string& err) {
    if (g_nccl.lib) return NMPM_OK;
    void* lib = nullptr;
    if (path && *path) lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already loaded by the host process
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return NMPM_ERR_INVALID;
    }
#define NMPM_NCCL_SYM(field, name)                                          \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name)); \
    if (!g_nccl.field) {                                                    \
        err = std::string("libnccl: missing symbol ") + name;               \
        return NMPM_ERR_INVALID;                                            \
    }
    NMPM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NMPM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NMPM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NMPM_NCCL_SYM(Send, "ncclSend")
    NMPM_NCCL_SYM(Recv, "ncclRecv")
    NMPM_NCCL_SYM(AllGather, "ncclAllGather")
    NMPM_NCCL_SYM(GroupStart, "ncclGroupStart")
    NMPM_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NMPM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NMPM_NCCL_SYM
    g_nccl.lib = lib;
    return NMPM_OK;
}

constexpr int kTableInts = 12;  // per rank: n_left, n_right, n_kept, overflow, box lo[3], box hi[3], pad[2]

struct Rect {
    int x_plane, a0, na, b0, nb;
    size_t nodes() const { return (size_t) 2 * na * nb; }
};

}  // namespace

struct nmpm_slab_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    std::vector<int> bounds;       // ownership boundaries (world+1), as set by the caller
    std::vector<int> grid_bounds;  // boundaries the particles currently obey (lag one G2P behind `bounds`)
    size_t cap_records = 0, rec_words = 0, plane_nodes = 0;
    float4 *pl_send[2] = {nullptr, nullptr}, *pl_recv[2] = {nullptr, nullptr};  // [0] = left, [1] = right neighbour
    float *mig_send[2] = {nullptr, nullptr}, *mig_recv[2] = {nullptr, nullptr};
    int *d_mine = nullptr, *d_table = nullptr, *h_table = nullptr;
    std::vector<int> boxes;  // world x 6 (lo[3], hi[3]) of the particles of the coming step; empty = unknown
    long long migrated = 0;
    cudaEvent_t ev_table = nullptr;
    bool p2g_issued = false;  // the P2G of the coming step is already on the stream (see slab_exchange_migrants)
    // NMPM_SLAB_TRACE=1: device time of the segments of a free-running step (events, one sync per nmpm_slab_step call)
    bool trace = false;
    std::vector<cudaEvent_t> tev;  // 5 per step: start, after P2G, after plane exchange, after G2P, after migrants
    double tsum[4] = {0, 0, 0, 0};
    long long tsteps = 0;
    int tskip = 8;
};

#define NCCL_TRY(h, expr)                                                                                   \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess) {                                                                            \
            (h)->last_error = std::string(#expr " failed: ") + g_nccl.GetErrorString(_r);                  \
            return NMPM_ERR_CUDA;                                                                           \
        }                                                                                                   \
    } while (0)

static void slab_comm_free(nmpm_sim* h) {
    nmpm_slab_comm* c = h->sc;
    if (!c) return;
    for (int s = 0; s < 2; ++s) {
        cudaFree(c->pl_send[s]), cudaFree(c->pl_recv[s]), cudaFree(c->mig_send[s]), cudaFree(c->mig_recv[s]);
    }
    cudaFree(c->d_mine), cudaFree(c->d_table);
    if (c->ev_table) cudaEventDestroy(c->ev_table);
    if (c->trace && c->tsteps)
        std::fprintf(stderr,
                     "[nmpm slab trace] rank %d: %lld steps, device ms/step: p2g(+sort,clear) %.3f | plane exchange %.3f | "
                     "grid_op+g2p %.3f | table+migrants %.3f\n",
                     c->rank, c->tsteps, c->tsum[0] / c->tsteps, c->tsum[1] / c->tsteps, c->tsum[2] / c->tsteps,
                     c->tsum[3] / c->tsteps);
    for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
    if (c->h_table) cudaFreeHost(c->h_table);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    h->sc = nullptr;
}

// rectangle (in-plane) that covers what this rank and its neighbour `nb` can write into their two shared planes
static Rect shared_rect(const nmpm_sim* h, int x_plane, int lo_rank, int hi_rank) {
    const nmpm_slab_comm* c = h->sc;
    const int n1 = h->res + 1;
    Rect r{x_plane, 0, h->dim == 3 ? n1 : 1, 0, n1};
    if (c->boxes.empty()) return r;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int) 0x80000000, (int) 0x80000000, (int) 0x80000000};
    for (int k = std::max(lo_rank, 0); k <= std::min(hi_rank, c->world - 1); ++k)
        for (int d = 0; d < 3; ++d) {
            lo[d] = std::min(lo[d], c->boxes[(size_t) k * 6 + d]);
            hi[d] = std::max(hi[d], c->boxes[(size_t) k * 6 + 3 + d]);
        }
    auto span = [&](int d, int* first, int* count) {
        if (hi[d] < lo[d]) {  // no particles anywhere near: nothing to exchange
            *first = 0, *count = 0;
            return;
        }
        // one cell of margin: the table describes the positions before this step's migrants were unpacked
        const int l = std::max(lo[d] - 1, 0), u = std::min(hi[d] + 3, n1 - 1);
        *first = l, *count = u - l + 1;
    };
    if (h->dim == 3) {
        span(1, &r.a0, &r.na);
        span(2, &r.b0, &r.nb);
    } else {
        span(1, &r.b0, &r.nb);
    }
    return r;
}

static int slab_exchange_planes(nmpm_sim* h) {
    nmpm_slab_comm* c = h->sc;
    const int n1 = h->res + 1;
    const int nbr[2] = {c->rank > 0 ? c->rank - 1 : -1, c->rank < c->world - 1 ? c->rank + 1 : -1};
    Rect rect[2];
    // shared planes start at the boundary between the two ranks; the rectangle depends on ranks r-1..r+2 of the pair
    rect[0] = shared_rect(h, c->grid_bounds[c->rank], c->rank - 2, c->rank + 1);
    rect[1] = shared_rect(h, c->grid_bounds[c->rank + 1], c->rank - 1, c->rank + 2);
    for (int s = 0; s < 2; ++s) {
        if (nbr[s] < 0 || rect[s].nodes() == 0) continue;
        if (rect[s].x_plane + 2 > n1) return NMPM_ERR_INVALID;
        k_rect<0><<<blocks_for(rect[s].nodes(), 256), 256, 0, h->stream>>>(h->grid, c->plane_nodes, n1, rect[s].x_plane, 2,
                                                                           rect[s].a0, rect[s].na, rect[s].b0, rect[s].nb,
                                                                           c->pl_send[s]);
        h->launches++;
    }
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int s = 0; s < 2; ++s) {
        if (nbr[s] < 0 || rect[s].nodes() == 0) continue;
        NCCL_TRY(h, g_nccl.Send(c->pl_send[s], rect[s].nodes() * 4, ncclFloat32, nbr[s], c->comm, h->stream));
        NCCL_TRY(h, g_nccl.Recv(c->pl_recv[s], rect[s].nodes() * 4, ncclFloat32, nbr[s], c->comm, h->stream));
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    for (int s = 0; s < 2; ++s) {
        if (nbr[s] < 0 || rect[s].nodes() == 0) continue;
        k_rect<1><<<blocks_for(rect[s].nodes(), 256), 256, 0, h->stream>>>(h->grid, c->plane_nodes, n1, rect[s].x_plane, 2,
                                                                           rect[s].a0, rect[s].na, rect[s].b0, rect[s].nb,
                                                                           c->pl_recv[s]);
        h->launches++;
        h->dirty_rects.push_back({rect[s].x_plane, rect[s].a0, rect[s].na, rect[s].b0, rect[s].nb});
    }
    CUDA_TRY(h, cudaGetLastError());
    return NMPM_OK;
}

// `early_p2g`: the coming step is an in-place one inside the same nmpm_slab_step call; its P2G over the resident slots
// is issued BEFORE the host waits for the table, so the GPU works through it while the host sizes and posts the
// migrant exchange; the received particles are scattered afterwards (P2G is additive).
static int slab_exchange_migrants(nmpm_sim* h, bool early_p2g) {
    nmpm_slab_comm* c = h->sc;
    // counts (written by the G2P) + the node box of the coming step, gathered from every rank
    CUDA_TRY(h, cudaMemcpyAsync(c->d_mine + 4, h->d_box + h->box_cur, 6 * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    NCCL_TRY(h, g_nccl.AllGather(c->d_mine, c->d_table, kTableInts, ncclInt32, c->comm, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(c->h_table, c->d_table, (size_t) c->world * kTableInts * sizeof(int), cudaMemcpyDeviceToHost,
                                h->stream));
    CUDA_TRY(h, cudaEventRecord(c->ev_table, h->stream));
    const size_t first_new = h->n_store;
    if (early_p2g)
        if (int rc = slab_p2g_early(h)) return rc;
    CUDA_TRY(h, cudaEventSynchronize(c->ev_table));  // the one host wait of the step (table + error flag have landed)
    if (int rc = poll_error(h)) return rc;
    const int* t = c->h_table;
    c->boxes.resize((size_t) c->world * 6);
    for (int k = 0; k < c->world; ++k) {
        if (t[k * kTableInts + 3]) {
            h->last_error = "slab migration buffer overflow on rank " + std::to_string(k) + " (raise cap_records)";
            return NMPM_ERR_INVALID;
        }
        for (int d = 0; d < 6; ++d) c->boxes[(size_t) k * 6 + d] = t[k * kTableInts + 4 + d];
    }
    const int r = c->rank;
    const size_t n_send[2] = {(size_t) t[r * kTableInts + 0], (size_t) t[r * kTableInts + 1]};
    const int nbr[2] = {r > 0 ? r - 1 : -1, r < c->world - 1 ? r + 1 : -1};
    const size_t n_recv[2] = {nbr[0] >= 0 ? (size_t) t[nbr[0] * kTableInts + 1] : 0, nbr[1] >= 0 ? (size_t) t[nbr[1] * kTableInts + 0] : 0};
    if ((nbr[0] < 0 && n_send[0]) || (nbr[1] < 0 && n_send[1])) {
        h->last_error = "a particle left the outermost slab";
        return NMPM_ERR_INVALID;
    }
    if (n_recv[0] > c->cap_records || n_recv[1] > c->cap_records) {
        h->last_error = "slab migration receive buffer too small (raise cap_records)";
        return NMPM_ERR_INVALID;
    }
    if (n_send[0] + n_send[1] + n_recv[0] + n_recv[1]) {
        NCCL_TRY(h, g_nccl.GroupStart());
        for (int s = 0; s < 2; ++s) {
            if (n_send[s]) NCCL_TRY(h, g_nccl.Send(c->mig_send[s], n_send[s] * c->rec_words, ncclFloat32, nbr[s], c->comm, h->stream));
            if (n_recv[s]) NCCL_TRY(h, g_nccl.Recv(c->mig_recv[s], n_recv[s] * c->rec_words, ncclFloat32, nbr[s], c->comm, h->stream));
        }
        NCCL_TRY(h, g_nccl.GroupEnd());
    }
    c->migrated += (long long) (n_send[0] + n_send[1]);
    if (int rc = nmpm_slab_unpack(h, c->mig_recv[0], n_recv[0], c->mig_recv[1], n_recv[1], n_send[0] + n_send[1])) return rc;
    if (early_p2g) {
        if (int rc = slab_p2g_tail(h, first_new)) return rc;
        c->p2g_issued = true;
    }
    return NMPM_OK;
}

extern "C" {

int nmpm_nccl_unique_id(void* out128, const char* libnccl_path) {
    if (!out128) return NMPM_ERR_INVALID;
    if (int rc = nccl_load(libnccl_path, g_create_error)) return rc;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) {
        g_create_error = "ncclGetUniqueId failed";
        return NMPM_ERR_CUDA;
    }
    std::memcpy(out128, &id, sizeof(id));
    return NMPM_OK;
}

int nmpm_slab_comm_init(nmpm_handle h, const void* unique_id128, int rank, int world, const int* bounds, size_t cap_records,
                        const char* libnccl_path) {
    if (int rc = slab_check(h, "nmpm_slab_comm_init")) return rc;
    if (!unique_id128 || !bounds || world < 1 || rank < 0 || rank >= world || cap_records == 0 || h->sc) return NMPM_ERR_INVALID;
    if (int rc = nccl_load(libnccl_path, h->last_error)) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    nmpm_slab_comm* c = new nmpm_slab_comm();
    h->sc = c;
    c->rank = rank, c->world = world;
    c->bounds.assign(bounds, bounds + world + 1);
    c->grid_bounds = c->bounds;
    c->cap_records = cap_records;
    c->rec_words = nmpm_migrate_record_bytes(h) / sizeof(float);
    c->plane_nodes = nmpm_grid_plane_bytes(h) / sizeof(float4);
    ncclUniqueId id;
    std::memcpy(&id, unique_id128, sizeof(id));
    NCCL_TRY(h, g_nccl.CommInitRank(&c->comm, world, id, rank));
    for (int s = 0; s < 2; ++s) {
        CUDA_TRY(h, cudaMalloc(&c->pl_send[s], 2 * c->plane_nodes * sizeof(float4)));
        CUDA_TRY(h, cudaMalloc(&c->pl_recv[s], 2 * c->plane_nodes * sizeof(float4)));
        CUDA_TRY(h, cudaMalloc(&c->mig_send[s], cap_records * c->rec_words * sizeof(float)));
        CUDA_TRY(h, cudaMalloc(&c->mig_recv[s], cap_records * c->rec_words * sizeof(float)));
    }
    CUDA_TRY(h, cudaMalloc(&c->d_mine, kTableInts * sizeof(int)));
    CUDA_TRY(h, cudaMemset(c->d_mine, 0, kTableInts * sizeof(int)));
    CUDA_TRY(h, cudaMalloc(&c->d_table, (size_t) world * kTableInts * sizeof(int)));
    CUDA_TRY(h, cudaMallocHost(&c->h_table, (size_t) world * kTableInts * sizeof(int)));
    CUDA_TRY(h, cudaEventCreateWithFlags(&c->ev_table, cudaEventDisableTiming));
    const char* tr = std::getenv("NMPM_SLAB_TRACE");
    c->trace = tr && *tr && *tr != '0';
    return NMPM_OK;
}

int nmpm_slab_set_bounds(nmpm_handle h, const int* bounds) {
    if (int rc = slab_check(h, "nmpm_slab_set_bounds")) return rc;
    if (!h->sc || !bounds) return NMPM_ERR_INVALID;
    nmpm_slab_comm* c = h->sc;
    c->bounds.assign(bounds, bounds + c->world + 1);
    const int x0 = c->rank > 0 ? bounds[c->rank] : 0;
    const int x1 = c->rank < c->world - 1 ? bounds[c->rank + 1] : bounds[c->world] + (1 << 20);
    return nmpm_slab_set_range(h, x0, x1);
}

long long nmpm_slab_migrated(nmpm_handle h) { return (h && h->sc) ? h->sc->migrated : 0; }

int nmpm_slab_step(nmpm_handle h, int nsteps) {
    if (int rc = slab_check(h, "nmpm_slab_step")) return rc;
    if (!h->sc || nsteps < 0) {
        h->last_error = "nmpm_slab_step: call nmpm_slab_comm_init first";
        return NMPM_ERR_INVALID;
    }
    nmpm_slab_comm* c = h->sc;
    if (c->trace) {
        while (c->tev.size() < (size_t) 5 * nsteps) {
            cudaEvent_t e;
            CUDA_TRY(h, cudaEventCreate(&e));
            c->tev.push_back(e);
        }
    }
    auto mark = [&](int s, int k) {
        if (c->trace) cudaEventRecord(c->tev[(size_t) 5 * s + k], h->stream);
    };
    for (int s = 0; s < nsteps; ++s) {
        mark(s, 0);
        if (c->p2g_issued) {
            c->p2g_issued = false;
        } else if (int rc = nmpm_slab_p2g(h)) {
            return rc;
        }
        mark(s, 1);
        if (int rc = slab_exchange_planes(h)) return rc;
        mark(s, 2);
        if (int rc = nmpm_slab_grid_g2p(h, c->mig_send[0], c->mig_send[1], c->cap_records, c->d_mine)) return rc;
        c->grid_bounds = c->bounds;  // after this G2P every particle obeys the current boundaries
        mark(s, 3);
        const bool early = !c->trace && s + 1 < nsteps && slab_next_step_in_place(h);  // the trace keeps segments apart
        if (int rc = slab_exchange_migrants(h, early)) return rc;
        mark(s, 4);
    }
    if (c->trace) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (int s = 0; s < nsteps; ++s) {
            if (c->tskip > 0) {  // communicator set-up and first-touch costs
                --c->tskip;
                continue;
            }
            ++c->tsteps;
            float ms = 0;
            const cudaEvent_t* e = &c->tev[(size_t) 5 * s];
            cudaEventElapsedTime(&ms, e[0], e[1]);
            c->tsum[0] += ms;
            cudaEventElapsedTime(&ms, e[1], e[2]);
            c->tsum[1] += ms;
            cudaEventElapsedTime(&ms, e[2], e[3]);
            c->tsum[2] += ms;
            cudaEventElapsedTime(&ms, e[3], e[4]);
            c->tsum[3] += ms;
        }
    }
    return NMPM_OK;
}

}  // extern "C"
