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
    size_t cap_records = 0, rec_words = 0, plane_nodes = 0;
    float4 *pl_send[2] = {nullptr, nullptr}, *pl_recv[2] = {nullptr, nullptr};  // [0] = left, [1] = right neighbour
    float *mig_send[2] = {nullptr, nullptr}, *mig_recv[2] = {nullptr, nullptr};  // (1 + cap_records) records each
    int *d_mine = nullptr, *d_ring = nullptr, *h_ring = nullptr;
    cudaEvent_t ev_ring[kRing] = {};
    long long step_no = 0;          // steps issued through nmpm_slab_step
    long long bounds_step = -1000;  // step at which the ownership boundaries last changed
    size_t k_hist[kRing] = {};      // records received at most in step s (sum of both message capacities)
    long long migrated = 0;
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

// capacity (records) of the migrant message exchanged with neighbour `side` in the coming step
static size_t slab_message_capacity(const nmpm_sim* h, const int* rec, int side) {
    const nmpm_slab_comm* c = h->sc;
    const bool has = side ? c->rank < c->world - 1 : c->rank > 0;
    if (!has) return 0;
    // no record yet, or the boundaries moved within the last few steps (whole planes of particles change owner at once)
    if (!rec || c->step_no - c->bounds_step < 2 * kLag + 2) return c->cap_records;
    const size_t sent = (size_t) rec[side], got = (size_t) rec[12 + 8 * side];
    return std::min(c->cap_records, 2 * std::max(sent, got) + 16384);
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
    const size_t msg_bytes = (1 + cap_records) * c->rec_words * sizeof(float);
    for (int s = 0; s < 2; ++s) {
        CUDA_TRY(h, cudaMalloc(&c->pl_send[s], 2 * c->plane_nodes * sizeof(float4)));
        CUDA_TRY(h, cudaMalloc(&c->pl_recv[s], 2 * c->plane_nodes * sizeof(float4)));
        CUDA_TRY(h, cudaMalloc(&c->mig_send[s], msg_bytes));
        CUDA_TRY(h, cudaMalloc(&c->mig_recv[s], msg_bytes));
        CUDA_TRY(h, cudaMemset(c->mig_send[s], 0, c->rec_words * sizeof(float)));
        CUDA_TRY(h, cudaMemset(c->mig_recv[s], 0, c->rec_words * sizeof(float)));
    }
    CUDA_TRY(h, cudaMalloc(&c->d_mine, 12 * sizeof(int)));
    CUDA_TRY(h, cudaMemset(c->d_mine, 0, 12 * sizeof(int)));
    CUDA_TRY(h, cudaMalloc(&c->d_ring, kRingInts * sizeof(int)));
    CUDA_TRY(h, cudaMemset(c->d_ring, 0, kRingInts * sizeof(int)));
    CUDA_TRY(h, cudaMallocHost(&c->h_ring, (size_t) kRing * kRingInts * sizeof(int)));
    std::memset(c->h_ring, 0, (size_t) kRing * kRingInts * sizeof(int));
    for (cudaEvent_t& e : c->ev_ring) CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->dev_counts = true;
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

long long nmpm_slab_migrated(nmpm_handle h) { return (h && h->sc) ? h->sc->migrated : 0; }

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
                h->last_error = "slab step: migration send buffer overflow (raise cap_records)";
                return NMPM_ERR_INVALID;
            }
            if ((!has[0] && rec[0]) || (!has[1] && rec[1])) {
                h->last_error = "a particle left the outermost slab";
                return NMPM_ERR_INVALID;
            }
            c->migrated += (long long) rec[0] + rec[1];
            // slots in use: true count after step e, plus whatever the unpacks since then may have appended
            size_t bound = (size_t) rec[28];
            for (long long k = e + 1; k < c->step_no; ++k) bound += c->k_hist[k % kRing];
            h->n_store = std::min(h->cap, std::max(bound, (size_t) 1));
        }
        const size_t K[2] = {slab_message_capacity(h, rec, 0), slab_message_capacity(h, rec, 1)};
        const bool stable = c->step_no - c->bounds_step >= 2 * kLag + 2;

        mark(s, 0);
        if (int rc = nmpm_slab_p2g(h)) return rc;
        mark(s, 1);
        if (int rc = slab_exchange_planes(h, stable ? rec : nullptr)) return rc;
        mark(s, 2);
        if (int rc = slab_grid_g2p2(h, c->mig_send[0] + W, K[0], c->mig_send[1] + W, K[1], c->d_mine)) return rc;
        c->grid_bounds = c->bounds;  // after this G2P every particle obeys the current boundaries
        mark(s, 3);
        // ---- migrants: headers, one NCCL group, append behind the device-side slot counter ------------------------------
        NMPM_DISPATCH_DIM(h, (k_slab_headers<D><<<1, 32, 0, h->stream>>>(c->d_mine, h->d_box + h->box_cur, (int*) c->mig_send[0],
                                                                        (int*) c->mig_send[1], h->d_ctr, (int) c->step_no,
                                                                        c->d_ring)));
        h->launches++;
        if (has[0] || has[1]) {
            NCCL_TRY(h, g_nccl.GroupStart());
            for (int sd = 0; sd < 2; ++sd) {
                if (!has[sd]) continue;
                const int nbr = sd ? c->rank + 1 : c->rank - 1;
                NCCL_TRY(h, g_nccl.Send(c->mig_send[sd], (1 + K[sd]) * W, ncclFloat32, nbr, c->comm, h->stream));
                NCCL_TRY(h, g_nccl.Recv(c->mig_recv[sd], (1 + K[sd]) * W, ncclFloat32, nbr, c->comm, h->stream));
            }
            NCCL_TRY(h, g_nccl.GroupEnd());
        }
        const float* rl = has[0] ? c->mig_recv[0] : nullptr;
        const float* rr = has[1] ? c->mig_recv[1] : nullptr;
        if (K[0] + K[1]) {
            NMPM_DISPATCH_DIM(h, (k_unpack_records2<D><<<blocks_for(K[0] + K[1], 256), 256, 0, h->stream>>>(
                                     rl, rr, (uint32_t) K[0], (uint32_t) K[1], (uint32_t) h->cap, h->store[h->cur], h->P,
                                     h->tiles_per_axis, h->sort.keys_a, h->d_box + h->box_cur, h->d_ctr)));
            h->launches++;
        }
        k_ctr_after_unpack<<<1, 32, 0, h->stream>>>(rl, rr, (uint32_t) K[0], (uint32_t) K[1], (uint32_t) h->cap, h->d_ctr, c->d_ring);
        h->launches++;
        CUDA_TRY(h, cudaMemcpyAsync(c->h_ring + (size_t) (c->step_no % kRing) * kRingInts, c->d_ring, kRingInts * sizeof(int),
                                    cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaEventRecord(c->ev_ring[c->step_no % kRing], h->stream));
        c->k_hist[c->step_no % kRing] = K[0] + K[1];
        h->n_store = std::min(h->cap, h->n_store + K[0] + K[1]);
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
    }
    if (particles) *particles = np;
    if (slots_in_use) *slots_in_use = ns;
    return NMPM_OK;
}

}  // extern "C"
