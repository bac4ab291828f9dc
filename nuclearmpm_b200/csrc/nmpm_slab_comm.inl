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
typedef enum { ncclInt8 = 0, ncclInt32 = 2, ncclFloat32 = 7 } ncclDataType_t;
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

constexpr int kRing = 4;        // read-back records kept on the host (the step uses the record of two steps ago)
constexpr int kRingInts = 32;   // [0..3] counts, [4..9] own box, [10] step | [12..19] header from the left | [20..27] from the right | [28..31] ctr
constexpr int kLag = 2;

struct Rect {
    int x_plane, a0, na, b0, nb;
    size_t nodes() const { return (size_t) 2 * na * nb; }
};

}  // namespace

// Device-driven slab step.  Nothing the host needs to size a launch or a message comes from the step in flight:
//   * particle counts live on the device (nmpm_sim::d_ctr); the host keeps an upper bound of the slots in use and every
//     slot beyond the true count, like every migrated-away slot, carries kKeyGone in the key array and is skipped;
//   * migrants travel in fixed-capacity messages [header | records]: the header carries the count (and the sender's node
//     box), the receiver appends behind its device-side slot counter;
//   * message capacity and the in-plane rectangle of the ghost exchange are derived from the read-back record of TWO
//     steps ago (own counts/box + the two received headers, copied to pinned memory after every step) — both ends of a
//     pair hold the same numbers, so they size the exchange identically without talking;
// so the host enqueues steps up to two ahead of the GPU and launch latency disappears behind the kernels; the only
// blocking call is the wait for a record that is two steps old.
struct nmpm_slab_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    std::vector<int> bounds;       // ownership boundaries (world+1), as set by the caller
    std::vector<int> grid_bounds;  // boundaries the particles currently obey (lag one G2P behind `bounds`)
    size_t cap_records = 0, rec_words = 0, plane_nodes = 0, n_global = 0;
    float4 *pl_send[2] = {nullptr, nullptr}, *pl_recv[2] = {nullptr, nullptr};  // [0] = left, [1] = right neighbour
    // Migrants travel through peer memory (CUDA IPC over NVLink): `inbox` holds this rank's four receive buffers
    // [from side 0/1][step parity], (1 + cap_records) records each; G2P writes the records of a leaving particle straight
    // into the NEIGHBOUR's inbox, k_slab_post adds header + arrival flag, k_slab_wait on the other side waits for it.
    // No message size has to be agreed and only the bytes of actual migrants cross the link.
    float* inbox = nullptr;
    int* flags = nullptr;                          // [0] posted by the left neighbour, [1] by the right one
    float* peer_inbox[2] = {nullptr, nullptr};     // the left / right neighbour's `inbox`
    int* peer_flags[2] = {nullptr, nullptr};
    float* dummy = nullptr;                        // record sink for a side without neighbour (capacity 0)
    size_t inbox_stride = 0;                       // floats per receive buffer
    int *d_mine = nullptr, *d_ring = nullptr, *h_ring = nullptr;
    cudaEvent_t ev_ring[kRing] = {};
    long long step_no = 0;          // steps issued through nmpm_slab_step
    long long bounds_step = -1000;  // step at which the ownership boundaries last changed
    long long migrated = 0, acct_step = 0;  // records of steps < acct_step are in `migrated`
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
    cudaStreamSynchronize(h->stream);
    for (int s = 0; s < 2; ++s) {
        cudaFree(c->pl_send[s]), cudaFree(c->pl_recv[s]);
        if (c->peer_inbox[s]) cudaIpcCloseMemHandle(c->peer_inbox[s]);
        if (c->peer_flags[s]) cudaIpcCloseMemHandle(c->peer_flags[s]);
    }
    cudaFree(c->inbox), cudaFree(c->flags), cudaFree(c->dummy);
    cudaFree(c->d_mine), cudaFree(c->d_ring);
    for (cudaEvent_t e : c->ev_ring)
        if (e) cudaEventDestroy(e);
    if (c->trace && c->tsteps)
        std::fprintf(stderr,
                     "[nmpm slab trace] rank %d: %lld steps, device ms/step: p2g(+sort,clear) %.3f | plane exchange %.3f | "
                     "grid_op+g2p %.3f | migrants %.3f\n",
                     c->rank, c->tsteps, c->tsum[0] / c->tsteps, c->tsum[1] / c->tsteps, c->tsum[2] / c->tsteps,
                     c->tsum[3] / c->tsteps);
    for (cudaEvent_t e : c->tev) cudaEventDestroy(e);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
    h->sc = nullptr;
}

// read-back record of step `s` (complete: its event has been waited for), or null when there is none yet
static const int* slab_record(nmpm_sim* h, long long s) {
    nmpm_slab_comm* c = h->sc;
    if (s < 0 || s + kRing <= c->step_no) return nullptr;
    return c->h_ring + (size_t) (s % kRing) * kRingInts;
}

// In-plane rectangle of the two node planes shared with neighbour `side`: everything this rank's and that neighbour's
// particles can touch there during the coming P2G.  `rec` = read-back record of two steps ago (null: whole planes).
static Rect shared_rect(const nmpm_sim* h, int x_plane, const int* rec, int side) {
    const int n1 = h->res + 1;
    Rect r{x_plane, 0, h->dim == 3 ? n1 : 1, 0, n1};
    if (!rec) return r;
    const int* mine = rec + 4;                 // lo[3], hi[3]
    const int* theirs = rec + 12 + 8 * side + 1;
    int lo[3], hi[3];
    for (int d = 0; d < 3; ++d) {
        lo[d] = std::min(mine[d], theirs[d]);
        hi[d] = std::max(mine[3 + d], theirs[3 + d]);
    }
    auto span = [&](int d, int* first, int* count) {
        if (hi[d] < lo[d]) {  // no particles on either side: nothing to exchange
            *first = 0, *count = 0;
            return;
        }
        // the boxes describe the positions two G2Ps ago: < 1 cell of motion per step (grid velocity clamp) + 1 of slack
        const int l = std::max(lo[d] - (kLag + 1), 0), u = std::min(hi[d] + 2 + (kLag + 1), n1 - 1);
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

static int slab_exchange_planes(nmpm_sim* h, const int* rec) {
    nmpm_slab_comm* c = h->sc;
    const int n1 = h->res + 1;
    const int nbr[2] = {c->rank > 0 ? c->rank - 1 : -1, c->rank < c->world - 1 ? c->rank + 1 : -1};
    Rect rect[2];
    rect[0] = shared_rect(h, c->grid_bounds[c->rank], rec, 0);
    rect[1] = shared_rect(h, c->grid_bounds[c->rank + 1], rec, 1);
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

// the step's G2P half with per-side send capacities (nmpm_slab_grid_g2p is the public single-capacity form)
static int slab_grid_g2p2(nmpm_sim* h, float* send_left, size_t cap_left, float* send_right, size_t cap_right, int* d_counts) {
    if (h->phase_next != 1) {
        h->last_error = "slab step: P2G must come first";
        return NMPM_ERR_INVALID;
    }
    CUDA_TRY(h, cudaMemsetAsync(d_counts, 0, 4 * sizeof(int), h->stream));
    if (h->timing) cudaEventRecord(h->ev[5], h->stream);
    if (int rc = do_grid_op(h)) return rc;
    if (h->timing) cudaEventRecord(h->ev[3], h->stream);
    MigrateArgs mig{h->opt.slab_x0, h->opt.slab_x1, send_left, send_right, (uint32_t) cap_left, d_counts, (uint32_t) cap_right};
    // a side without neighbour has capacity 0: a particle leaving the outermost slab raises the overflow flag
    if (int rc = do_g2p(h, mig)) return rc;
    if (h->timing) {
        cudaEventRecord(h->ev[4], h->stream);
        cudaEventSynchronize(h->ev[4]);
        float ms = 0;
        cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
        h->t_ms[NMPM_T_SORT] += ms;
        cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]);
        h->t_ms[NMPM_T_P2G] += ms;
        cudaEventElapsedTime(&ms, h->ev[5], h->ev[3]);
        h->t_ms[NMPM_T_GRID] += ms;
        cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]);
        h->t_ms[NMPM_T_G2P] += ms;
        h->t_steps++;
    }
    h->steps_done++;
    h->phase_next = 0;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(h->h_error, h->d_error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
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
    if (h->steps_done != 0 || h->n_gone != 0) {
        h->last_error = "nmpm_slab_comm_init: only right after creation";
        return NMPM_ERR_INVALID;
    }
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
    }
    c->inbox_stride = (1 + cap_records) * c->rec_words;
    CUDA_TRY(h, cudaMalloc(&c->inbox, 4 * c->inbox_stride * sizeof(float)));
    CUDA_TRY(h, cudaMemset(c->inbox, 0, 4 * c->inbox_stride * sizeof(float)));
    CUDA_TRY(h, cudaMalloc(&c->flags, 64));
    CUDA_TRY(h, cudaMemset(c->flags, 0, 64));
    CUDA_TRY(h, cudaMalloc(&c->dummy, 2 * c->rec_words * sizeof(float)));
    CUDA_TRY(h, cudaMemset(c->dummy, 0, 2 * c->rec_words * sizeof(float)));
    {   // exchange the IPC handles of inbox + flags (one all-gather at set-up), open the two neighbours'
        struct Handles {
            cudaIpcMemHandle_t inbox, flags;
        } mine_h;
        CUDA_TRY(h, cudaIpcGetMemHandle(&mine_h.inbox, c->inbox));
        CUDA_TRY(h, cudaIpcGetMemHandle(&mine_h.flags, c->flags));
        char *d_in = nullptr, *d_all = nullptr;
        CUDA_TRY(h, cudaMalloc(&d_in, sizeof(Handles)));
        CUDA_TRY(h, cudaMalloc(&d_all, (size_t) world * sizeof(Handles)));
        CUDA_TRY(h, cudaMemcpyAsync(d_in, &mine_h, sizeof(Handles), cudaMemcpyHostToDevice, h->stream));
        NCCL_TRY(h, g_nccl.AllGather(d_in, d_all, sizeof(Handles), ncclInt8, c->comm, h->stream));
        std::vector<Handles> all((size_t) world);
        CUDA_TRY(h, cudaMemcpyAsync(all.data(), d_all, (size_t) world * sizeof(Handles), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        cudaFree(d_in), cudaFree(d_all);
        for (int sd = 0; sd < 2; ++sd) {
            const int nb = sd ? rank + 1 : rank - 1;
            if (nb < 0 || nb >= world) continue;
            CUDA_TRY(h, cudaIpcOpenMemHandle((void**) &c->peer_inbox[sd], all[(size_t) nb].inbox, cudaIpcMemLazyEnablePeerAccess));
            CUDA_TRY(h, cudaIpcOpenMemHandle((void**) &c->peer_flags[sd], all[(size_t) nb].flags, cudaIpcMemLazyEnablePeerAccess));
        }
    }
    CUDA_TRY(h, cudaMalloc(&c->d_mine, 12 * sizeof(int)));
    CUDA_TRY(h, cudaMemset(c->d_mine, 0, 12 * sizeof(int)));
    CUDA_TRY(h, cudaMalloc(&c->d_ring, kRingInts * sizeof(int)));
    CUDA_TRY(h, cudaMemset(c->d_ring, 0, kRingInts * sizeof(int)));
    CUDA_TRY(h, cudaMallocHost(&c->h_ring, (size_t) kRing * kRingInts * sizeof(int)));
    std::memset(c->h_ring, 0, (size_t) kRing * kRingInts * sizeof(int));
    for (cudaEvent_t& e : c->ev_ring) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->dev_counts = true;
    c->n_global = (size_t) world * h->n;  // estimate until nmpm_slab_set_global_count is called
    const char* tr = std::getenv("NMPM_SLAB_TRACE");
    c->trace = tr && *tr && *tr != '0';
    return NMPM_OK;
}

int nmpm_slab_set_bounds(nmpm_handle h, const int* bounds) {
    if (int rc = slab_check(h, "nmpm_slab_set_bounds")) return rc;
    if (!h->sc || !bounds) return NMPM_ERR_INVALID;
    nmpm_slab_comm* c = h->sc;
    c->bounds.assign(bounds, bounds + c->world + 1);
    c->bounds_step = c->step_no;
    const int x0 = c->rank > 0 ? bounds[c->rank] : 0;
    const int x1 = c->rank < c->world - 1 ? bounds[c->rank + 1] : bounds[c->world] + (1 << 20);
    return nmpm_slab_set_range(h, x0, x1);
}

// records sent so far (synchronises the stream: the step only accounts records that are two steps old)
long long nmpm_slab_migrated(nmpm_handle h) {
    if (!h || !h->sc) return 0;
    nmpm_slab_comm* c = h->sc;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (; c->acct_step < c->step_no; ++c->acct_step) {
        const int* r2 = slab_record(h, c->acct_step);
        if (r2) c->migrated += (long long) r2[0] + r2[1];
    }
    return c->migrated;
}

int nmpm_slab_step(nmpm_handle h, int nsteps) {
    if (int rc = slab_check(h, "nmpm_slab_step")) return rc;
    if (!h->sc || nsteps < 0) {
        h->last_error = "nmpm_slab_step: call nmpm_slab_comm_init first";
        return NMPM_ERR_INVALID;
    }
    nmpm_slab_comm* c = h->sc;
    CUDA_TRY(h, cudaSetDevice(h->device));
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
    const size_t W = c->rec_words;
    const bool has[2] = {c->rank > 0, c->rank < c->world - 1};
    for (int s = 0; s < nsteps; ++s) {
        // ---- the record of two steps ago: errors, message capacities, exchange rectangle, slot bound --------------
        const int* rec = nullptr;
        if (c->step_no >= kLag) {
            const long long e = c->step_no - kLag;
            CUDA_TRY(h, cudaEventSynchronize(c->ev_ring[e % kRing]));  // two steps old: the host does not stall here
            rec = slab_record(h, e);
            if (int rc = poll_error(h)) return rc;
            if (rec[28 + 2] & 1) {
                h->last_error = "slab step: particle capacity exceeded (raise nmpm_options.capacity)";
                return NMPM_ERR_INVALID;
            }
            if ((rec[28 + 2] & 2) || rec[3]) {
                h->last_error = "slab step: migration buffer overflow (raise cap_records), or a particle left the outermost slab";
                return NMPM_ERR_INVALID;
            }
            if (rec[28 + 2] & 8) {
                h->last_error = "slab step: more slots in use than the launches covered (particle burst beyond the growth bound)";
                return NMPM_ERR_INVALID;
            }
            if (rec[28 + 2] & 4) {
                h->last_error = "slab step: a neighbour never posted its migrants (peer flag wait timed out)";
                return NMPM_ERR_CUDA;
            }
            if ((!has[0] && rec[0]) || (!has[1] && rec[1])) {
                h->last_error = "a particle left the outermost slab";
                return NMPM_ERR_INVALID;
            }
            for (; c->acct_step <= e; ++c->acct_step) {
                const int* r2 = slab_record(h, c->acct_step);
                if (r2) c->migrated += (long long) r2[0] + r2[1];
            }
        }
        // Slots the launches of this step cover.  The host does not know how many are in use: the true count of two steps
        // ago plus generous growth (one unpack has happened since: two sides, each at most twice what arrived then plus
        // 3 % of the store for a burst that starts from nothing), capped by the allocation.  k_slab_post verifies the bound
        // on the device (error bit 8).  Slots beyond the true count carry kKeyGone and are skipped, so a loose bound only
        // costs empty warps and sort keys (8 ranks on cfg4: about +20 % slots).
        if (c->step_no >= 1) {
            size_t bound = h->cap;
            if (rec) {
                const size_t n_true = (size_t) rec[28], arrived = (size_t) std::max(rec[12], rec[20]);
                const size_t growth = 2 * arrived + n_true / 32 + 16384;  // per side
                bound = std::min(h->cap, n_true + 2 * growth);
            }
            h->n_store = bound;
        }
        const bool stable = c->step_no - c->bounds_step >= 2 * kLag + 2;
        const int step_tag = (int) c->step_no + 1;                 // what the flags carry (0 = nothing posted yet)
        const size_t par = (size_t) (c->step_no & 1);
        // where this step's migrants go: the neighbours' inboxes [their side facing me][parity]; none: a sink of capacity 0
        float* out[2];
        size_t K[2];
        for (int sd = 0; sd < 2; ++sd) {
            out[sd] = has[sd] ? c->peer_inbox[sd] + ((size_t) (1 - sd) * 2 + par) * c->inbox_stride : c->dummy;
            K[sd] = has[sd] ? c->cap_records : 0;
        }
        const float* in[2] = {has[0] ? c->inbox + (0 * 2 + par) * c->inbox_stride : nullptr,
                              has[1] ? c->inbox + (1 * 2 + par) * c->inbox_stride : nullptr};

        mark(s, 0);
        if (int rc = nmpm_slab_p2g(h)) return rc;
        mark(s, 1);
        if (int rc = slab_exchange_planes(h, stable ? rec : nullptr)) return rc;
        mark(s, 2);
        if (int rc = slab_grid_g2p2(h, out[0] + W, K[0], out[1] + W, K[1], c->d_mine)) return rc;
        c->grid_bounds = c->bounds;  // after this G2P every particle obeys the current boundaries
        mark(s, 3);
        // ---- migrants: header + flag into the neighbours' inboxes, wait for theirs, append behind the device-side counter ----
        NMPM_DISPATCH_DIM(h, (k_slab_post<D><<<1, 32, 0, h->stream>>>(c->d_mine, h->d_box + h->box_cur, has[0] ? (int*) out[0] : nullptr,
                                                                     has[1] ? (int*) out[1] : nullptr, h->d_ctr, step_tag, c->d_ring,
                                                                     (int) std::min(h->n_store, (size_t) 0x7fffffff),
                                                                     has[0] ? c->peer_flags[0] + 1 : nullptr,
                                                                     has[1] ? c->peer_flags[1] + 0 : nullptr)));
        k_slab_wait<<<1, 32, 0, h->stream>>>(has[0] ? c->flags + 0 : nullptr, has[1] ? c->flags + 1 : nullptr, step_tag,
                                             (int*) in[0], (int*) in[1], h->d_ctr);
        h->launches += 2;
        if (has[0] || has[1]) {
            // launch size: the two-step-old counts with head-room; the kernel strides, so any count is handled
            size_t guess = 65536;
            if (rec) guess = 2 * (size_t) std::max(rec[12], rec[20]) + 2 * (size_t) std::max(rec[0], rec[1]) + 65536;
            guess = std::min(guess, 2 * c->cap_records);
            NMPM_DISPATCH_DIM(h, (k_unpack_records2<D><<<blocks_for(guess, 256), 256, 0, h->stream>>>(
                                     in[0], in[1], (uint32_t) c->cap_records, (uint32_t) c->cap_records, (uint32_t) h->cap,
                                     h->store[h->cur], h->P, h->tiles_per_axis, h->sort.keys_a, h->d_box + h->box_cur, h->d_ctr)));
            h->launches++;
        }
        k_ctr_after_unpack<<<1, 32, 0, h->stream>>>(in[0], in[1], (uint32_t) c->cap_records, (uint32_t) c->cap_records,
                                                    (uint32_t) h->cap, h->d_ctr, c->d_ring);
        h->launches++;
        CUDA_TRY(h, cudaMemcpyAsync(c->h_ring + (size_t) (c->step_no % kRing) * kRingInts, c->d_ring, kRingInts * sizeof(int),
                                    cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventRecord(c->ev_ring[c->step_no % kRing], h->stream));
        ++c->step_no;
        mark(s, 4);
        CUDA_TRY(h, cudaGetLastError());
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

// particle count of the whole (global) simulation: every rank must pass the same number (it sizes the migrant messages)
int nmpm_slab_set_global_count(nmpm_handle h, size_t n_global) {
    if (int rc = slab_check(h, "nmpm_slab_set_global_count")) return rc;
    if (!h->sc) return NMPM_ERR_INVALID;
    h->sc->n_global = n_global;
    return NMPM_OK;
}

// true particle / slot counts of a device-driven slab (synchronises the stream)
int nmpm_slab_counts(nmpm_handle h, long long* particles, long long* slots_in_use) {
    if (int rc = slab_check(h, "nmpm_slab_counts")) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    long long np = (long long) (h->n_store - h->n_gone), ns = (long long) h->n_store;
    if (h->dev_counts) {
        int ctr[4];
        CUDA_TRY(h, cudaMemcpyAsync(ctr, h->d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        np = (long long) ctr[0] - ctr[1], ns = ctr[0];
        if (ctr[2]) {  // the step itself reports these two steps late
            h->last_error = std::string("slab step error bits ") + std::to_string(ctr[2]) +
                            " (1: particle capacity exceeded, 2: migration buffer overflow / particle left the outermost slab, "
                            "4: peer flag wait timed out, 8: slot bound exceeded)";
            if (particles) *particles = np;
            if (slots_in_use) *slots_in_use = ns;
            return NMPM_ERR_INVALID;
        }
    }
    if (particles) *particles = np;
    if (slots_in_use) *slots_in_use = ns;
    return NMPM_OK;
}

}  // extern "C"
